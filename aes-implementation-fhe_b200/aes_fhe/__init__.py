"""Host-side mirror of the reference's AES-on-CKKS caller stack."""
from .context import EngineContext
from .pipeline import (AESPipeline, BatchedStateEncoder, CapturedRound, FipsDriver, RowMajorShiftRows, decrypt_readme_order)
from .steps import (AddRoundKey, InvMixColumnsFHE, InvShiftRows, MixColFinal, ShiftRows, StateEncoder, SubBytesLUT,
                    XOR4LUT, from_zeta, to_zeta)
from .snap import (NoiseReducer, Zeta16NoiseReducer, Zeta16Snap, Zeta16Snap1D, Zeta16SnapNoMul, Zeta16SnapPair,
                   load_coeff1d)
from .tables import expand_aes128_key, load_all_coeffs

__all__ = [n for n in dir() if not n.startswith("_")]
