"""Regenerate tests/golden/aes_reference_golden.json from the UNCHANGED reference modules.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference's .py files are imported as they are (R0 shims of SURVEY.md Appendix C via
tests/refload.py) on the tier-A slot-domain stand-in (oracle/slot_standin.py), because the
reference's own backend (`desilofhe`, closed wheel, unpinned) is not installable here.
Recorded: per-tag `_log_pair` bytes of the as-shipped encrypt/decrypt, per-primitive engine-op
counts, and sha256 digests of the full engine-call trace (at 256 slots, to keep it small).
"""
import json
import os
import sys
from collections import Counter
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
os.environ.setdefault("STANDIN_SLOTS", "256")

import numpy as np  # noqa: E402

import refload  # noqa: E402
from oracle import slot_standin as ss  # noqa: E402

FIPS_KEY = bytes(range(16))
FIPS_PT = bytes.fromhex("00112233445566778899aabbccddeeff")


def key_schedule(ref_test_mod, key):
    return ref_test_mod.expand_aes128_key(np.frombuffer(key, dtype=np.uint8).copy())


def load_test_driver(ref):
    """The reference key schedule lives in its __main__-style test script; import it as a module."""
    import importlib.util
    for m, mod in vars(ref).items():
        if hasattr(mod, "__name__"):
            sys.modules[mod.__name__] = mod
    spec = importlib.util.spec_from_file_location("ref_roundtrip", refload.REF / "test" / "test_aes_pipeline_roundtrip.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build(ref, coeffs):
    ctx = ref.engine_context.EngineContext(1, mode="cpu", thread_count=4)
    ctx.engine.trace_enabled = True
    x4 = ref.xor4_lut.XOR4LUT(ctx, coeffs["xor4"])
    pipe = ref.pipeline.AESPipeline(ctx, coeffs, mixcolumns=ref.mixcol_final.MixColFinal(ctx, x4),
                                    inv_mixcolumns=ref.invmixcolumns_fhe.InvMixColumnsFHE(ctx, x4),
                                    use_hard_renorm_between_steps=True)
    return ctx, pipe


def counts(eng):
    c = Counter()
    for (stage, op), n in eng.counters.items():
        c[op] += n
    return dict(sorted(c.items()))


def main():
    ref = refload.load(ss)
    drv = load_test_driver(ref)
    coeffs = drv.load_all_coeffs(ref.coeff_dir)
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes

    np.random.seed(7)
    k7 = np.random.randint(0, 256, 16, dtype=np.uint8)
    p7 = np.random.randint(0, 256, 16, dtype=np.uint8)
    cases = {"fips_c1": (FIPS_KEY, FIPS_PT), "seed7": (bytes(k7), bytes(p7))}
    out = {"slots": int(os.environ["STANDIN_SLOTS"]), "cases": {}, "primitives": {}}

    ctx, pipe = build(ref, coeffs)
    out["construct"] = {"digest": ctx.engine.trace_digest(), "ops": counts(ctx.engine)}
    for name, (key, pt) in cases.items():
        rks = key_schedule(drv, key)
        aes = Cipher(algorithms.AES(key), modes.ECB()).encryptor()
        case = {"key": key.hex(), "pt": pt.hex(), "fips_ct": (aes.update(pt) + aes.finalize()).hex(),
                "round_keys": [bytes(r).hex() for r in rks]}
        ctx.engine.reset_trace()
        pipe._rk_cache = None
        dbg = {}
        ct = pipe.encrypt(np.frombuffer(pt, dtype=np.uint8).copy(), rks, dbg)
        case["enc_tags"] = {t: bytes(e["plain"]).hex() for t, e in dbg.items()}
        case["enc_digest"] = ctx.engine.trace_digest()
        case["enc_ops"] = counts(ctx.engine)
        ctx.engine.reset_trace()
        dbg = {}
        dec = pipe.decrypt(*ct, rks, dbg)
        case["dec_tags"] = {t: bytes(e["plain"]).hex() for t, e in dbg.items()}
        case["dec_digest"] = ctx.engine.trace_digest()
        case["dec_ops"] = counts(ctx.engine)
        out["cases"][name] = case

    # per-primitive traces on a seed-0 state (as the reference self-tests use, mixcol_final.py:264-265)
    rng = np.random.RandomState(0)
    st = rng.randint(0, 256, 16, dtype=np.uint8)
    ky = rng.randint(0, 256, 16, dtype=np.uint8)
    prims = {
        "add_round_key": lambda c, k: pipe.add_round_key(*c, *k),
        "sub_bytes": lambda c, k: pipe.sub_bytes(*c),
        "inv_sub_bytes": lambda c, k: pipe.inv_sub_bytes(*c),
        "shift_rows": lambda c, k: pipe.shift_rows(*c),
        "inv_shift_rows": lambda c, k: pipe.inv_shift_rows(*c),
        "mix_columns": lambda c, k: pipe.mix_columns(*c),
        "inv_mix_columns": lambda c, k: pipe.inv_mix_columns(*c),
    }
    for name, fn in prims.items():
        ctx.engine.reset_trace()
        c = pipe.encoder.encode(st.copy())
        k = pipe.encoder.encode(ky.copy())
        res = fn(c, k)
        dig, ops = ctx.engine.trace_digest(), counts(ctx.engine)
        out["primitives"][name] = {"state": bytes(st).hex(), "key": bytes(ky).hex(),
                                   "out": bytes(pipe.encoder.decode(*res)).hex(), "digest": dig, "ops": ops}
    (HERE / "aes_reference_golden.json").write_text(json.dumps(out, indent=1, sort_keys=True))
    print("wrote", HERE / "aes_reference_golden.json")


if __name__ == "__main__":
    main()
