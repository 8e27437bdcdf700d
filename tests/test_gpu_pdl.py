"""Programmatic dependent launch against plain stream order. Almost every kernel of the library is launched with
programmatic stream serialization: it may start while its predecessor still runs and waits at griddepcontrol.wait. An SM's
L1 is not coherent, and under these launches a line from an EARLIER call has been seen to be served again to a plain load
(that is how a bad tile table was once read) — so everything a predecessor wrote, or a caller rewrites between
calls, must be read through L2. This test replays a history of calls (front end with device and host points, the inference
network, whole training steps — different inputs every call, every buffer reused) twice in fresh processes, with
LISEC_NO_PDL=1 (plain stream order, the reference) and without, and requires identical bits from every call."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_worker(no_pdl: bool) -> dict:
    env = dict(os.environ)
    env["LISEC_NO_PDL"] = "1" if no_pdl else "0"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "pdl_worker.py")], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("PDL_WORKER ")][-1]
    return json.loads(line[len("PDL_WORKER "):])


def test_overlapped_launches_produce_the_bits_of_plain_stream_order():
    plain, pdl = run_worker(True), run_worker(False)
    assert set(plain) == set(pdl) == {"frontend_device", "frontend_host", "network", "train"}
    for key in plain:
        assert plain[key] == pdl[key], (key, [i for i, (a, b) in enumerate(zip(plain[key], pdl[key])) if a != b])
    # and the history really exercised different results per call
    assert len(set(plain["frontend_device"])) >= 3 and len(set(plain["train"])) == len(plain["train"])
