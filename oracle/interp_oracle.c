/*
 * TEST INFRASTRUCTURE ONLY (oracle).  See interp_oracle_impl.h.
 * Built by oracle/Makefile into oracle/liboracle_interp.so; loaded by
 * oracle/nufft_oracle.py through ctypes.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it.
 */
#include <math.h>
#include <string.h>
#include <stddef.h>

#define REAL float
#define NAME(x) orc_f32_##x
#include "interp_oracle_impl.h"
#undef REAL
#undef NAME

#define REAL double
#define NAME(x) orc_f64_##x
#include "interp_oracle_impl.h"
#undef REAL
#undef NAME
