// Stands where the reference's src/internal.h stands on the include path of oracle/ref_shim.cpp -- the TU that drives
// the bridge functions in the reference's own order (the call sequence of src/visodo.cpp:1041-1415 and
// src/keyframe_align.cpp:178-350).  Compiling that TU UNCHANGED against this header and linking librgbid_b200.so is
// the compiler's proof that include/rgbid_b200/internal.hpp is source compatible with src/internal.h for every call
// the hot path makes (tests/test_refloop_gpu.py compares its results with the same TU built on the reference's kernels).
#pragma once
#include "rgbid_b200/internal.hpp"
