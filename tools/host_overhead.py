"""Host enqueue time vs device time of engine calls (is the stream starved by the host?)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "aes-implementation-fhe_b200")]
import numpy as np
import desilofhe

eng = desilofhe.Engine(logn=16, levels=21, use_bootstrap=True, seed=1)
sk = eng.create_secret_key(); eng.create_public_key(sk); rk = eng.create_relinearization_key(sk)
eng.create_conjugation_key(sk); bk = eng.create_bootstrap_key(sk)
z = np.exp(2j * np.pi * np.random.default_rng(0).random(eng.slot_count))
a = eng.encrypt(z); b = eng.encrypt(z)
def measure(name, fn, reps=5):
    fn(); eng.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()                       # results are dropped as a real flow does (steady-state arena behaviour)
    t1 = time.perf_counter()
    eng.sync()
    t2 = time.perf_counter()
    c = eng.counters()
    print(f"{name:12s} host enqueue {1e3*(t1-t0)/reps:8.3f} ms/call   total {1e3*(t2-t0)/reps:8.3f} ms/call")
l0 = eng.counters()["launches"]
measure("mul", lambda: eng.multiply(a, b, rk), 20)
print("launches per mul", (eng.counters()["launches"] - l0) / 21)
measure("rotate", lambda: eng.rotate(a, None, 5), 20)
measure("add", lambda: eng.add(a, b), 50)
measure("mul_const", lambda: eng.multiply(a, 0.5), 20)
l0 = eng.counters()["launches"]
measure("bootstrap", lambda: eng.bootstrap(a), 3)
print("launches per bootstrap", (eng.counters()["launches"] - l0) / 4)
measure("boot pair", lambda: eng.pair_map(eng.bootstrap, (a,), (b,)), 3)
print(eng.arena_stats())
