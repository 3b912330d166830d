"""Timeline of block 0 of the fused kernel (DBX_FUSED_TRACE=1): per-event clock deltas."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DBX_FUSED_TRACE"] = "1"
import deepbayes_b200 as db
from deepbayes_b200 import _lib
from deepbayes_b200.engine import Engine
from oracle import oracle as o

dims = [784, 512, 512, 10]
L = o.mlp(dims)
m = db.Sequential([db.Dense(512, activation="relu", input_shape=(784,)), db.Dense(512, activation="relu"), db.Dense(10, activation="softmax")])
eng = Engine(m)
mean, std = o.synthetic_posterior(L, seed=1)
eng.set_gaussian(mean, std)
x = o.synthetic_x((4096, 784), seed=0)
for _ in range(2):
    eng.predict_sums(x, 256, seed=1, mode="bf16")
lib = _lib.load()
buf = (C.c_longlong * 16000)()
lib.dbx_debug_trace.restype = C.c_int
n = lib.dbx_debug_trace(buf, 8000)
ev = np.array(buf[:2 * n]).reshape(n, 2)
ev = ev[np.argsort(ev[:, 1], kind="stable")]
t0 = ev[0, 1]
names = {150: "  M l1 stage wait", 151: "  M l1 stage ready", 160: "  M l2 stage wait", 161: "  M l2 stage ready", 100: "M L1a start", 101: "M L1b start", 110: "M L1a issued", 111: "M L1b issued", 140: "M L3 issue",
         200: "E acc1[0] seen", 201: "E acc1[1] seen", 210: "E pack1[0] done", 211: "E pack1[1] done"}
for j in range(4):
    names[120 + j] = "M L2c%d start" % j; names[130 + j] = "M L2c%d issued" % j
    names[220 + j] = "E acc2 c%d seen" % j; names[230 + j] = "E pack2 c%d done" % j; names[240 + j] = "E l3 c%d drained" % j
# print units 10..12 (steady state)
starts = [i for i in range(n) if ev[i, 0] == 100]
print("events", n, "units", len(starts))
if os.environ.get("DBX_FUSED_TRACE_ONLY"):
    ids = sorted(set(int(v) for v in ev[:, 0]))
    # median offset of every selected probe from the preceding probe 100 of the same unit
    for pid in ids:
        if pid == 100:
            continue
        offs = []
        for a, b in zip(starts[5:-1], starts[6:]):
            for i in range(a + 1, b):
                if ev[i, 0] == pid:
                    offs.append(ev[i, 1] - ev[a, 1])
                    break
        if offs:
            print("probe %d: median offset from unit start %.0f" % (pid, np.median(offs)))
    d = np.diff(ev[starts, 1])
    print("cycles per unit with a single probe per unit: median %.0f  mean %.0f  min %d  max %d" % (np.median(d), d.mean(), d.min(), d.max()))
    sys.exit(0)
if len(starts) > 13:
    a, b = starts[10], starts[11]
    base = ev[a, 1]
    compact = os.environ.get("TRACE_COMPACT", "1") == "1"
    wait_acc = 0
    for i in range(a, b):
        if compact and int(ev[i, 0]) in (150, 160):
            continue
        if compact and int(ev[i, 0]) in (151, 161):
            wait_acc += ev[i, 1] - ev[i - 1, 1]
            continue
        if wait_acc:
            print("                    (stage waits since last line: %d cycles)" % wait_acc)
            wait_acc = 0
        print("%8d  +%6d  %s" % (ev[i, 1] - base, ev[i, 1] - ev[i - 1, 1], names.get(int(ev[i, 0]), str(ev[i, 0]))))
    per = (ev[starts[-1], 1] - ev[starts[5], 1]) / (len(starts) - 1 - 5)
    print("cycles per unit (steady):", per)
