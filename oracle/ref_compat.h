/* ref_compat.h -- force-included (nvcc -include) when the UNMODIFIED reference sources are
 * compiled with CUDA 12.9 / g++ 13 for oracle/_ref.  TEST INFRASTRUCTURE ONLY.
 *
 * /root/reference/src/SmpcController.cu:1620 and :1660 call abs() on a size_t difference, which is
 * ambiguous with today's <cstdlib> overload set (the reference targeted CUDA 7 / gcc 4.8).  An exact
 * overload for unsigned long resolves the call without touching the source; the value semantics
 * (|a-b| of an unsigned difference is the difference itself) are what the old toolchain computed.
 */
#pragma once
#include <cstddef>
#include <cstdlib>
#include <cmath>
static inline unsigned long abs(unsigned long v) { return v; }
