// cuda_host_shim.h -- TEST INFRASTRUCTURE.  Lets g++ compile numbakit-ode_b200/csrc/nbk_device.cuh
// and nbk_program.cu for the HOST so that the kernel logic (queue refill, retire, guards,
// dense output, start-up) can be exercised by `pytest -m "not gpu"` without a GPU.
// A "warp" has ONE lane here (NBK_FULL_MASK == 1); nothing in the product links this.
#pragma once
#include <cmath>
#include <cstring>

#define NBK_HOST_EMUL 1
#define NBK_FULL_MASK 1u
#define NBK_REFILL_MIN 1
#define __device__
#define __host__
#define __forceinline__ inline
#define __global__
#define __constant__
#define __grid_constant__
#define __launch_bounds__(...)

struct emul_dim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local emul_dim3 threadIdx, blockIdx, blockDim;

static inline int __double2hiint(double d) { long long v; std::memcpy(&v, &d, 8); return (int)(v >> 32); }
static inline int __double2loint(double d) { long long v; std::memcpy(&v, &d, 8); return (int)(v & 0xffffffffLL); }
static inline double __hiloint2double(int hi, int lo) { unsigned long long v = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo; double d; std::memcpy(&d, &v, 8); return d; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
template <class T> static inline T __ldg(const T *p) { return *p; }
using std::fabs; using std::fma; using std::fmax; using std::fmin; using std::pow; using std::sqrt;
