"""Back-of-envelope timeline model of one direction of the persistent recurrence (rnn_tc.cu), built from the phase
durations measured on the B200 (DESIGN.md section 5, step-100 timeline, cycles at 1.965 GHz).  It reproduces the measured
step (~10.5 k cycles) and evaluates the next-round candidates of DESIGN.md section 10 under the same assumptions:

  python scripts/recurrence_timeline.py

Model: every CTA of a direction runs the same schedule, so one CTA stands for all.  Per batch group and step:
  publish(t) -> barrier opens after BARRIER -> fence + TMA issue ISSUE -> first K-chunk group lands after LAND
  -> MMA phase: G groups, each REGION + n_mma * MMA cycles, a group can only start once it has landed; the ring holds
     RING groups, a load is issued when its slot is free and lands LAND later
  -> epilogue EPI_PRE (tcgen05.ld + gate math + staging + h stores) -> publish -> (y stores, off the critical path)
The tensor pipe and the epilogue warps are each a single resource shared by the groups in flight.
"""
import argparse

BARRIER, ISSUE, LAND = 1950, 300, 1100          # publish -> barrier open; fence + TMA issue; issue -> group landed
REGION, MMA_64, MMA_128 = 200, 46, 58           # elected region; 64x64x16 and 128x64x16 (SS) per instruction
EPI_PRE, EPI_POST = 2050, 650                   # epilogue up to the publish; release (store acks) before the next poll


def simulate(groups_in_flight=1, n_groups=5, mma_per_group=16, mma_cycles=MMA_64, ring=2, extra_epilogue=0, steps=40):
    """Returns cycles per (one step of every group in flight)."""
    pipe_free = 0.0                               # tensor pipe
    epi_free = 0.0                                # epilogue warps
    slot_free = [0.0] * ring                      # ring slots (time their previous content was consumed)
    publish = [0.0] * groups_in_flight            # time h_t of the group became visible to the release
    slot_i = 0
    marks = []
    for s in range(steps):
        for g in range(groups_in_flight):
            ready = publish[g] + BARRIER + ISSUE  # producer may issue this group's loads from here on
            t = pipe_free
            for k in range(n_groups):
                issue = max(ready, slot_free[slot_i])
                landed = issue + LAND
                start = max(t, landed)
                t = start + REGION + mma_per_group * mma_cycles
                slot_free[slot_i] = t
                slot_i = (slot_i + 1) % ring
            pipe_free = t
            e0 = max(t, epi_free)
            publish[g] = e0 + EPI_PRE + extra_epilogue + EPI_POST
            epi_free = publish[g]
        marks.append(publish[-1])
    return (marks[-1] - marks[len(marks) // 2]) / (len(marks) - 1 - len(marks) // 2)


def main():
    argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter).parse_args()
    base = simulate()
    print("current kernel (1 group of 64 rows, 5 x 16 MMAs of 64x64, ring of 2):  %6.0f cycles / step   (measured ~10 500)" % base)
    ks = simulate(n_groups=4, mma_per_group=10, mma_cycles=MMA_128, extra_epilogue=700)
    print("K split over a CTA pair (4 x 10 MMAs of 128x64, +0.7 k exchange):      %6.0f cycles / step   (%.2f x)" % (ks, base / ks))
    il = simulate(groups_in_flight=2)
    print("two batch groups in flight, shared W_hh and ring:                      %6.0f cycles / 2 steps (%.2f x per sequence)" % (il, 2 * base / il))
    il3 = simulate(groups_in_flight=3)
    print("three batch groups in flight:                                          %6.0f cycles / 3 steps (%.2f x per sequence)" % (il3, 3 * base / il3))
    both = simulate(groups_in_flight=2, n_groups=4, mma_per_group=10, mma_cycles=MMA_128, extra_epilogue=700)
    print("both:                                                                  %6.0f cycles / 2 steps (%.2f x per sequence)" % (both, 2 * base / both))


if __name__ == "__main__":
    main()
