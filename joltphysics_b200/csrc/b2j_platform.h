// b2j_platform.h -- build-mode glue.
//
// The kernels of this library are written as "one work item" device functions plus thin __global__ wrappers. The same
// bodies can be compiled for the host by tests/hostsim (B2J_HOSTSIM): a debugging aid for this GPU-less container that lets
// kernel logic be stepped against the reference oracle before spending GPU time. The shipped library (libjolt_b200.so)
// is always compiled by nvcc for sm_100a WITHOUT B2J_HOSTSIM and has no host execution path.
#pragma once

#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <math.h>
#include <float.h>

#if defined(B2J_HOSTSIM)
	#define B2J_HD inline
	#define B2J_D inline
	#define B2J_FORCEINLINE inline
	#define B2J_RESTRICT __restrict__
#else
	#include <cuda_runtime.h>
	#define B2J_HD __host__ __device__ __forceinline__
	#define B2J_D __device__ __forceinline__
	#define B2J_FORCEINLINE __forceinline__
	#define B2J_RESTRICT __restrict__
#endif

namespace b2j {

// ---- atomics (device: hardware atomics; hostsim: serial execution, plain ops) ----------------------------------
#if defined(B2J_HOSTSIM)
B2J_D uint32_t atomic_add(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
B2J_D uint32_t atomic_or(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
B2J_D uint32_t atomic_and(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o & v; return o; }
B2J_D uint32_t atomic_min(uint32_t *p, uint32_t v) { uint32_t o = *p; if (v < o) *p = v; return o; }
B2J_D uint32_t atomic_max(uint32_t *p, uint32_t v) { uint32_t o = *p; if (v > o) *p = v; return o; }
B2J_D uint32_t atomic_cas(uint32_t *p, uint32_t cmp, uint32_t v) { uint32_t o = *p; if (o == cmp) *p = v; return o; }
B2J_D uint32_t atomic_exch(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = v; return o; }
B2J_D unsigned long long atomic_cas64(unsigned long long *p, unsigned long long cmp, unsigned long long v) { unsigned long long o = *p; if (o == cmp) *p = v; return o; }
B2J_D uint32_t volatile_load(const uint32_t *p) { return *p; }
B2J_D void prefetch_l2(const void *) { }
B2J_D void grid_dependency_sync() { }
B2J_D void atomic_add_matched(uint32_t *arr, uint32_t index, uint32_t v) { arr[index] += v; }
template <bool kLockstep> B2J_D bool warp_any(bool p) { return p; }
B2J_D void mem_fence() { }
B2J_D int ctz32(uint32_t v) { return v == 0? 32 : __builtin_ctz(v); }
B2J_D int clz32(uint32_t v) { return v == 0? 32 : __builtin_clz(v); }
B2J_D int clz64(uint64_t v) { return v == 0? 64 : __builtin_clzll(v); }
#else
B2J_D uint32_t atomic_add(uint32_t *p, uint32_t v) { return atomicAdd(p, v); }
B2J_D uint32_t atomic_or(uint32_t *p, uint32_t v) { return atomicOr(p, v); }
B2J_D uint32_t atomic_and(uint32_t *p, uint32_t v) { return atomicAnd(p, v); }
B2J_D uint32_t atomic_min(uint32_t *p, uint32_t v) { return atomicMin(p, v); }
B2J_D uint32_t atomic_max(uint32_t *p, uint32_t v) { return atomicMax(p, v); }
B2J_D uint32_t atomic_cas(uint32_t *p, uint32_t cmp, uint32_t v) { return atomicCAS(p, cmp, v); }
B2J_D uint32_t atomic_exch(uint32_t *p, uint32_t v) { return atomicExch(p, v); }
B2J_D unsigned long long atomic_cas64(unsigned long long *p, unsigned long long cmp, unsigned long long v) { return atomicCAS(p, cmp, v); }
B2J_D uint32_t volatile_load(const uint32_t *p) { return *(const volatile uint32_t *)p; }
B2J_D void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
// Programmatic dependent launch (Runtime::launch_pdl): everything before this call may run while the previous kernel of the stream is
// still running, so it may only touch data that kernel does not write. wait = the previous kernel has completed and its writes are
// visible; launch_dependents = the next kernel of the stream may start its own prologue now.
B2J_D void grid_dependency_sync()
{
	asm volatile("griddepcontrol.wait;" ::: "memory");
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
// arr[index] += v where many lanes of a warp are likely to hit the same index: lanes with equal index are combined into one atomic
B2J_D void atomic_add_matched(uint32_t *arr, uint32_t index, uint32_t v)
{
	uint32_t active = __activemask();
	uint32_t peers = __match_any_sync(active, index);
	uint32_t lane = threadIdx.x & 31u;
	// v is summed over the peers with shuffles only when it differs; the callers add small constants, so count * v is enough
	if ((uint32_t)(__ffs((int)peers) - 1) == lane)
		atomicAdd(&arr[index], v * (uint32_t)__popc(peers));
}
// vote over the whole warp when kLockstep (every lane of the warp must call it), identity otherwise
template <bool kLockstep> B2J_HD bool warp_any(bool p)
{
#if defined(__CUDA_ARCH__)
	if (kLockstep) return __any_sync(0xffffffffu, p) != 0;
#endif
	return p;
}
B2J_D void mem_fence() { __threadfence(); }
B2J_D int ctz32(uint32_t v) { return v == 0? 32 : __ffs((int)v) - 1; }
B2J_D int clz32(uint32_t v) { return __clz((int)v); }
B2J_D int clz64(uint64_t v) { return __clzll((long long)v); }
#endif

} // namespace b2j
