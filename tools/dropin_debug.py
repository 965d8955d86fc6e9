import sys, re, pathlib, subprocess, os
sys.path.insert(0, ".")
from tests.test_dropin import _capture, _run, CPU_BIN, GPU_BIN
tmp = pathlib.Path("/tmp/dd"); tmp.mkdir(exist_ok=True)
cap, nb = _capture(tmp, [-50_000, -175_000], nblk=40, seed=7, acars=True)
freqs = ["136.975", "136.850"]
print("bursts", nb)
def ids(txt):
    return sorted(re.findall(r'"hex":"(\w+)".*?NUMBER (\d+)', txt))
for rep in range(3):
    a = _run(CPU_BIN, cap, freqs, extra=("-J",))[0]
    b, berr = _run(GPU_BIN, cap, freqs, extra=("-J",))
    ia, ib = ids(a), ids(b)
    print(rep, "cpu", len(ia), "gpu", len(ib), "gpu missing", sorted(set(ia) - set(ib)), "cpu missing", sorted(set(ib) - set(ia)))
print(berr[-600:])
# which bursts does the oracle decode per channel (no threads, no races)?
import numpy as np
from oracle.pyoracle import Oracle
iq = np.fromfile(cap, dtype=np.uint8)
for fo in (-50_000, -175_000):
    o = Oracle("port", Fo=fo).feed(iq); oq = Oracle("port", Fo=fo).feed(iq, "rtl_quirk")
    print("oracle Fo", fo, "blocks", len(o.blocks), "with rtl quirk", len(oq.blocks))
