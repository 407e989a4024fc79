"""Transcript stitching of the streaming engines (host logic, no device work)."""


def stitch_transcript(iterating, transcript):
    """The reference's "collapsing characters hack" (DanSpeechRecognizer.py:169-174): when a chunk's greedy transcript
    starts with the character the running transcript ends with, that character is taken to be the same CTC emission
    cut in two by the chunk border and is dropped.  Returns (new running transcript, the part this chunk added)."""
    if iterating and transcript and iterating[-1] == transcript[0]:
        transcript = transcript[1:]
    return iterating + transcript, transcript
