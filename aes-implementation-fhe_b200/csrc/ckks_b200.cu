// ckks_b200.cu -- unity build of the engine: one translation unit so nvcc compiles (and inlines across)
// all kernels once.  Build: see __graft_entry__.build() (nvcc, sm_100a) and tests/emu/build.py (g++, CKKS_EMU).
#include "ntt.cu"
#include "kernels.cu"
#include "lut.cu"
#include "engine.cu"
#include "bootstrap.cu"
#include "capi.cu"
