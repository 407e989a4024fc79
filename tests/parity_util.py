"""Shared parity metrics for the model output (tests and the `parity` key of bench.py).

probs are [T, C] softmax rows of ONE utterance (valid frames only); `ref` is the oracle's output for the same input.
"""
import numpy as np

MARGIN_EDGES = (0.0, 1e-3, 1e-2, 3e-2, 1e-1, 3e-1, 1.0, 3.0, 10.0, np.inf)


def edit_distance(a, b):
    """Levenshtein distance of two strings (what decoder.py:45-74 computes through the absent C extension)."""
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def frame_report(p, ref, floor=1e-30):
    """Frame-level agreement of the greedy path and the oracle's argmax margin (top1 - top2 log-probability) at the
    frames where the path differs: a histogram over MARGIN_EDGES plus the largest margin that flipped."""
    p = np.asarray(p, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    a, r = p.argmax(1), ref.argmax(1)
    lr = np.log(np.maximum(ref, floor))
    srt = np.sort(lr, axis=1)
    margin = srt[:, -1] - srt[:, -2]
    diff = a != r
    hist = np.histogram(margin[diff], bins=np.asarray(MARGIN_EDGES))[0].tolist()
    all_hist = np.histogram(margin, bins=np.asarray(MARGIN_EDGES))[0].tolist()
    return {"frames": int(len(r)), "frames_differ": int(diff.sum()),
            "max_flipped_margin": float(margin[diff].max()) if diff.any() else 0.0,
            "flipped_margin_hist": hist, "margin_hist": all_hist}


def merge_reports(reps):
    out = {"frames": 0, "frames_differ": 0, "max_flipped_margin": 0.0,
           "flipped_margin_hist": [0] * (len(MARGIN_EDGES) - 1), "margin_hist": [0] * (len(MARGIN_EDGES) - 1)}
    for r in reps:
        out["frames"] += r["frames"]
        out["frames_differ"] += r["frames_differ"]
        out["max_flipped_margin"] = max(out["max_flipped_margin"], r["max_flipped_margin"])
        out["flipped_margin_hist"] = [x + y for x, y in zip(out["flipped_margin_hist"], r["flipped_margin_hist"])]
        out["margin_hist"] = [x + y for x, y in zip(out["margin_hist"], r["margin_hist"])]
    out["margin_edges"] = [float(e) if np.isfinite(e) else "inf" for e in MARGIN_EDGES]
    return out


def transcript_report(got, ref):
    """Identity rate and character error rate of greedy transcripts against the oracle's."""
    same = sum(int(a == b) for a, b in zip(got, ref))
    edits = sum(edit_distance(a, b) for a, b in zip(got, ref))
    chars = sum(len(b) for b in ref)
    return {"utterances": len(ref), "identical": same, "identical_rate": same / max(len(ref), 1),
            "char_edits": edits, "ref_chars": chars, "cer": edits / max(chars, 1)}
