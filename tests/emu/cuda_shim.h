/* TEST INFRASTRUCTURE -- not part of the product.
 *
 * Just enough of the CUDA device vocabulary for g++ to compile the engine's DEVICE LOGIC headers
 * (ra_b200/csrc/raft_step.cuh, raft_row.cuh) as plain host C++, one "thread" at a time, so that
 * the CPU test tier can diff that very logic against the oracle without a GPU (tests/emu/ra_emu.cpp).
 * Include every system header BEFORE this file: it defines qualifier-like macros. */
#pragma once
#define RA_HOST_EMU 1
#define CTA_T 1                      /* one thread per "CTA": the per-thread shared-memory columns are plain arrays */
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__

struct ulonglong2 { unsigned long long x, y; };
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { ulonglong2 r; r.x = x; r.y = y; return r; }
struct uint2 { unsigned int x, y; };
static inline uint2 make_uint2(unsigned int x, unsigned int y) { uint2 r; r.x = x; r.y = y; return r; }

struct ra_emu_dim3 { unsigned x, y, z; };
static const ra_emu_dim3 threadIdx = {0, 0, 0};

static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
/* a "warp" of one lane */
static inline unsigned __activemask() { return 1u; }
template <typename T> static inline unsigned __match_any_sync(unsigned, T) { return 1u; }
template <typename T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline unsigned atomicCAS(unsigned* p, unsigned cmp, unsigned v) { unsigned o = *p; if (o == cmp) *p = v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
