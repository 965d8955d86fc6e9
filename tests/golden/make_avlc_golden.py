"""Generates tests/golden/avlc_golden.json: synthetic ACARS-over-AVLC frames and the JSON line the REFERENCE's out()
(out.c / outacars.c compiled in place, oracle/_ref/libvdl2outref.so) prints for each.  Run where /root/reference is mounted:
    python tests/golden/make_avlc_golden.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pyoracle
from tests.test_avlc_oracle import GOLD, _cases

if __name__ == "__main__":
    pyoracle.build("all")
    out = [{"frame": f.hex(), "json": pyoracle.out_json(f, t=1577836800.0).strip()} for f in _cases(n=28, seed=11)]
    assert all(o["json"].startswith("{") for o in out)
    json.dump(out, open(GOLD, "w"), indent=0)
    print(f"{len(out)} frames -> {GOLD}")
