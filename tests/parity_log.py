"""Parity evidence: every GPU parity assertion appends what it measured (achieved error, applied
tolerance, iteration counts) to gpurun_out/parity_log.jsonl; scripts/parity_table.py turns the log of
a GPU run into profiles/r2_parity_table.md (tracked)."""
import json
import os
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# one file per run (gpurun merges gpurun_out/ back file by file: a fixed name would be overwritten by the next call)
LOG = os.path.join(ROOT, "gpurun_out", f"parity_log_{os.environ.get('FDAL_PARITY_TAG') or time.strftime('%m%d_%H%M')}.jsonl")


def _plain(v):
    try:
        import numpy as np

        if isinstance(v, (np.floating, np.integer)):
            return v.item()
    except Exception:
        pass
    if isinstance(v, (tuple, list)):
        return [_plain(x) for x in v]
    if isinstance(v, dict):
        return {str(k): _plain(x) for k, x in v.items()}
    return v


def record(test: str, values: dict):
    try:
        os.makedirs(os.path.dirname(LOG), exist_ok=True)
        with open(LOG, "a") as f:
            f.write(json.dumps({"test": test, **_plain(values)}) + "\n")
    except Exception:
        pass


def check(what: str, err: float, tol: float, **extra):
    """Assert err < tol and log both (the test id comes from pytest's PYTEST_CURRENT_TEST)."""
    test = os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0].split("::", 1)[-1]
    record(test, {"what": what, "err": float(err), "tol": float(tol), **extra})
    assert err < tol, (what, err, tol)
