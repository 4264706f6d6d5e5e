"""`natsort.natsorted` (reference: diff_completion_pipeline.py:11,196 — scan files in natural order)."""
import re

_NUM = re.compile(r"(\d+)")


def natsort_key(s):
    return [int(t) if t.isdigit() else t.lower() for t in _NUM.split(str(s))]


def natsorted(seq, key=None, reverse=False):
    return sorted(seq, key=(lambda v: natsort_key(key(v))) if key else natsort_key, reverse=reverse)
