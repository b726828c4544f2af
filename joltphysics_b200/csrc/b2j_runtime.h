// b2j_runtime.h -- device memory, kernel launch and device-wide primitives (sort / scan).
//
// CUDA build: plain cudaMalloc'd arrays, grid-stride launches sized in multiples of the SM count, CUB radix sort / scan.
// B2J_HOSTSIM build (tests/hostsim only, never shipped): the same kernels run as serial loops, see b2j_platform.h.
#pragma once

#include "b2j_platform.h"

#include <algorithm>
#include <numeric>
#include <string>
#include <vector>
#include <mutex>
#include <map>
#include <stdio.h>
#include <stdlib.h>
#include <typeinfo>
#include <cxxabi.h>

#ifndef B2J_HOSTSIM
#include <cub/cub.cuh>
#endif

namespace b2j {

inline std::string &last_error() { static thread_local std::string e; return e; }

#ifndef B2J_HOSTSIM
#define B2J_CUDA_CHECK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { last_error() = std::string(#expr) + ": " + cudaGetErrorString(_e); return false; } } while (0)

template <class K> __global__ void __launch_bounds__(128) run_kernel(const K k, uint32_t n)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
		k(i);
}
// same with an explicit block size / minimum resident blocks (register budget) for kernels tuned by measurement
template <class K, int THREADS, int MINB> __global__ void __launch_bounds__(THREADS, MINB) run_kernel_cfg(const K k, uint32_t n)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
		k(i);
}
// n = min(*n_ptr, cap) - *begin_ptr (begin_ptr may be null); the functor receives indices relative to begin
template <class K> __global__ void __launch_bounds__(128) run_kernel_dev(const K k, const uint32_t *n_ptr, const uint32_t *begin_ptr, uint32_t cap)
{
	uint32_t n = *n_ptr;
	if (n > cap) n = cap;
	uint32_t begin = begin_ptr != nullptr? *begin_ptr : 0;
	n = n > begin? n - begin : 0;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
		k(i);
}
// Lockstep form: EVERY lane calls k.run(i, valid) every round so that the item body can keep the warp converged with votes
template <class K> __global__ void __launch_bounds__(128) run_kernel_dev_lockstep(const K k, const uint32_t *n_ptr, const uint32_t *begin_ptr, uint32_t cap)
{
	uint32_t n = *n_ptr;
	if (n > cap) n = cap;
	uint32_t begin = begin_ptr != nullptr? *begin_ptr : 0;
	n = n > begin? n - begin : 0;
	for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x)
	{
		uint32_t i = base + threadIdx.x;
		k.run(i, i < n);
		__syncwarp();
	}
}
// One WARP per work item for serial, scratch hungry item bodies (EPA): lane 0 runs the item with an S scratch block in shared
// memory (low latency instead of a global memory slot); slot = global warp id (ownership of any global side scratch).
template <class K, class S> __global__ void __launch_bounds__(256) run_kernel_warp_smem(const K k, const uint32_t *n_ptr, uint32_t cap)
{
	extern __shared__ __align__(16) unsigned char b2j_smem[];
	uint32_t n = *n_ptr;
	if (n > cap) n = cap;
	uint32_t warp_in_block = threadIdx.x >> 5, lane = threadIdx.x & 31, warps_per_block = blockDim.x >> 5;
	uint32_t slot = blockIdx.x * warps_per_block + warp_in_block;
	S *scratch = reinterpret_cast<S *>(b2j_smem) + warp_in_block;
	for (uint32_t i = slot; i < n; i += gridDim.x * warps_per_block)
	{
		if (lane == 0)
			k.run(i, true, slot, *scratch);
		__syncwarp();
	}
}
// One WARP per work item, all 32 lanes inside the item body (run_warp): S = the warp's big scratch block, H = its small hand-over area,
// both in shared memory.
template <class K, class S, class H> __global__ void __launch_bounds__(256) run_kernel_warp_coop(const K k, const uint32_t *n_ptr, uint32_t cap)
{
	extern __shared__ __align__(16) unsigned char b2j_smem[];
	uint32_t n = *n_ptr;
	if (n > cap) n = cap;
	uint32_t warp_in_block = threadIdx.x >> 5, warps_per_block = blockDim.x >> 5;
	uint32_t slot = blockIdx.x * warps_per_block + warp_in_block;
	S *scratch = reinterpret_cast<S *>(b2j_smem) + warp_in_block;
	H *hand = reinterpret_cast<H *>(b2j_smem + (size_t)warps_per_block * sizeof(S)) + warp_in_block;
	for (uint32_t i = slot; i < n; i += gridDim.x * warps_per_block)
	{
		k.run_warp(i, slot, *scratch, *hand);
		__syncwarp();
	}
}
// One THREAD per work item with a private S scratch block in LOCAL memory (EPA: 2 KB or 21 KB per lane): local memory is word
// interleaved across the lanes of a warp, so lockstep lanes touching the same field coalesce, and it is L1/L2 cached. Every lane
// calls run() every round (valid = the lane has an item): the item bodies keep the warp in lockstep with votes (warp_any<true>),
// otherwise lanes drift apart for good (measured: 1.8 active lanes per instruction, 4x slower). Measured against a warp-per-item
// form with the scratch in shared memory (instruction fetch bound: 77% stall_no_inst) and a shared memory thread-per-item form.
template <class K, class S> __global__ void __launch_bounds__(128) run_kernel_lane_local(const K k, const uint32_t *n_ptr, uint32_t cap)
{
	S scratch;
	uint32_t n = *n_ptr;
	if (n > cap) n = cap;
	uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x)
	{
		uint32_t i = base + threadIdx.x;
		k.run(i, i < n, slot, scratch);
		__syncwarp();
	}
}
#endif

// kernel categories for the optional per kernel timing (one id per kernel functor type, process wide)
inline std::vector<std::string> &profile_names() { static std::vector<std::string> n; return n; }
inline std::mutex &profile_mutex() { static std::mutex m; return m; } // the groups of a batch launch from several host threads
template <class K> inline int profile_category()
{
	static int id = [] {
		int status = 0;
		char *dm = abi::__cxa_demangle(typeid(K).name(), nullptr, nullptr, &status);
		std::string name = dm != nullptr? dm : typeid(K).name();
		free(dm);
		for (size_t p = name.find("b2j::"); p != std::string::npos; p = name.find("b2j::")) name.erase(p, 5);
		if (name.compare(0, 16, "KSolveVelocityT<") == 0) name = "KSolveVelocity"; // (both forms of the kernel are one category)
		std::lock_guard<std::mutex> lock(profile_mutex());
		profile_names().push_back(name);
		return (int)profile_names().size() - 1;
	}();
	return id;
}

struct Runtime
{
	int device = 0;
	bool profiling = false;
	std::vector<double> prof_ms;
	std::vector<uint32_t> prof_launches;
#ifndef B2J_HOSTSIM
	struct ProfEvent { cudaEvent_t a, b; int cat; };
	std::vector<ProfEvent> prof_pending, prof_free;
	void prof_begin(int cat)
	{
		ProfEvent e;
		if (!prof_free.empty()) { e = prof_free.back(); prof_free.pop_back(); }
		else { cudaEventCreate(&e.a); cudaEventCreate(&e.b); }
		e.cat = cat;
		cudaEventRecord(e.a, stream);
		prof_pending.push_back(e);
	}
	void prof_end() { cudaEventRecord(prof_pending.back().b, stream); }
#else
	void prof_begin(int) { }
	void prof_end() { }
#endif
	// accumulate the timings of all launches since the last call (the stream must be idle)
	void prof_collect()
	{
#ifndef B2J_HOSTSIM
		for (ProfEvent &e : prof_pending)
		{
			float ms = 0.0f;
			if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess)
			{
				if ((size_t)e.cat >= prof_ms.size()) { prof_ms.resize(e.cat + 1, 0.0); prof_launches.resize(e.cat + 1, 0); }
				prof_ms[e.cat] += ms;
				prof_launches[e.cat] += 1;
			}
			prof_free.push_back(e);
		}
		prof_pending.clear();
#endif
	}
	void prof_reset() { prof_collect(); prof_ms.assign(prof_ms.size(), 0.0); prof_launches.assign(prof_launches.size(), 0); }

	int num_sms = 148;
	std::map<const void *, bool> func_configured;   // kernels whose function attributes were set on this runtime's device
	std::map<const void *, int> func_blocks_per_sm; // occupancy of the cooperative kernels on this runtime's device
	uint32_t launches = 0;          // kernels launched since the last reset
	void *cub_temp = nullptr;
	size_t cub_temp_size = 0;
#ifndef B2J_HOSTSIM
	cudaStream_t stream = nullptr;
	void *pinned_small = nullptr;
	static constexpr size_t kPinnedSmall = 64 * 1024;
#endif

	bool init(int dev)
	{
		device = dev;
#ifndef B2J_HOSTSIM
		int count = 0;
		cudaError_t e = cudaGetDeviceCount(&count);
		if (e != cudaSuccess || count == 0)
		{
			last_error() = std::string("no CUDA device available (libjolt_b200 has no CPU fallback): ") + cudaGetErrorString(e);
			return false;
		}
		B2J_CUDA_CHECK(cudaSetDevice(dev));
		cudaDeviceProp prop;
		B2J_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
		num_sms = prop.multiProcessorCount;
		B2J_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
		B2J_CUDA_CHECK(cudaMallocHost(&pinned_small, kPinnedSmall)); // landing zone of the small per step readbacks (counters, offsets)
#endif
		return true;
	}

	void shutdown()
	{
#ifndef B2J_HOSTSIM
		if (cub_temp) cudaFree(cub_temp);
		if (pinned_small) cudaFreeHost(pinned_small);
		if (stage_dev) cudaFree(stage_dev);
		if (stage_host) cudaFreeHost(stage_host);
		for (ProfEvent &e : prof_free) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
		prof_free.clear();
		if (stream) cudaStreamDestroy(stream);
#else
		free(cub_temp); free(stage_dev); free(stage_host);
#endif
		stage_dev = stage_host = nullptr; stage_cap = 0;
		cub_temp = nullptr;
	}

	template <class T> T *alloc(size_t n, bool zero = true)
	{
		if (n == 0) n = 1;
		T *p = nullptr;
#ifndef B2J_HOSTSIM
		if (cudaMalloc((void **)&p, n * sizeof(T)) != cudaSuccess) { last_error() = "cudaMalloc failed"; return nullptr; }
		if (zero) cudaMemsetAsync(p, 0, n * sizeof(T), stream);
#else
		p = (T *)malloc(n * sizeof(T));
		if (zero) memset((void *)p, 0, n * sizeof(T));
#endif
		return p;
	}
	template <class T> void free_(T *&p)
	{
		if (p == nullptr) return;
#ifndef B2J_HOSTSIM
		cudaFree((void *)p);
#else
		free((void *)p);
#endif
		p = nullptr;
	}
	template <class T> void upload(T *dst, const T *src, size_t n)
	{
		if (n == 0) return;
#ifndef B2J_HOSTSIM
		cudaMemcpyAsync((void *)dst, (const void *)src, n * sizeof(T), cudaMemcpyHostToDevice, stream);
		cudaStreamSynchronize(stream); // src is usually pageable / temporary
#else
		memcpy((void *)dst, (const void *)src, n * sizeof(T));
#endif
	}
	template <class T> void download(T *dst, const T *src, size_t n)
	{
		if (n == 0) return;
#ifndef B2J_HOSTSIM
		if (pinned_small != nullptr && n * sizeof(T) <= kPinnedSmall)
		{
			// a copy into pageable memory goes through a driver staging buffer with extra synchronisation: land in pinned memory
			cudaMemcpyAsync(pinned_small, (const void *)src, n * sizeof(T), cudaMemcpyDeviceToHost, stream);
			cudaStreamSynchronize(stream);
			memcpy((void *)dst, pinned_small, n * sizeof(T));
			return;
		}
		cudaMemcpyAsync((void *)dst, (const void *)src, n * sizeof(T), cudaMemcpyDeviceToHost, stream);
		cudaStreamSynchronize(stream);
#else
		memcpy((void *)dst, (const void *)src, n * sizeof(T));
#endif
	}
	template <class T> void copy(T *dst, const T *src, size_t n)
	{
		if (n == 0) return;
#ifndef B2J_HOSTSIM
		cudaMemcpyAsync((void *)dst, (const void *)src, n * sizeof(T), cudaMemcpyDeviceToDevice, stream);
#else
		memcpy((void *)dst, (const void *)src, n * sizeof(T));
#endif
	}
	void memset_(void *p, int v, size_t bytes)
	{
		if (bytes == 0) return;
#ifndef B2J_HOSTSIM
		cudaMemsetAsync(p, v, bytes, stream);
#else
		memset(p, v, bytes);
#endif
	}
	void sync()
	{
#ifndef B2J_HOSTSIM
		cudaStreamSynchronize(stream);
#endif
	}
	bool check(const char *what)
	{
#ifndef B2J_HOSTSIM
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) { last_error() = std::string(what) + ": " + cudaGetErrorString(e); return false; }
#endif
		(void)what;
		return true;
	}

	uint32_t grid_for(uint32_t n, uint32_t block) const
	{
		uint32_t g = (n + block - 1) / block;
		uint32_t cap = (uint32_t)num_sms * 16;
		return g < 1? 1 : (g > cap? cap : g);
	}

	// host-known count
	template <class K> void launch(const K &k, uint32_t n)
	{
		if (n == 0) return;
		++launches;
#ifndef B2J_HOSTSIM
		if (profiling) prof_begin(profile_category<K>());
		run_kernel<K><<<grid_for(n, 128), 128, 0, stream>>>(k, n);
		if (profiling) prof_end();
#else
		for (uint32_t i = 0; i < n; ++i) k(i);
#endif
	}
	// Programmatic dependent launch: the kernel may start while the previous kernel of the stream is still running; K calls
	// grid_dependency_sync() before it touches anything that kernel writes (chains of small dependent launches: the solver phases)
	template <class K> void launch_pdl(const K &k, uint32_t n)
	{
		if (n == 0) return;
#ifndef B2J_HOSTSIM
		if (profiling) { K serial = k; serial.pdl = 0; launch(serial, n); return; } // (per kernel event timing wants the kernels apart)
		++launches;
		cudaLaunchConfig_t cfg = {};
		cfg.gridDim = dim3(grid_for(n, 128)); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
		cudaLaunchAttribute attr[1];
		attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		attr[0].val.programmaticStreamSerializationAllowed = 1;
		cfg.attrs = attr; cfg.numAttrs = 1;
		cudaLaunchKernelEx(&cfg, run_kernel<K>, k, n);
#else
		launch(k, n);
#endif
	}
	// launch_pdl with an explicit block size / minimum resident blocks (register budget)
	template <class K, int THREADS, int MINB> void launch_pdl_cfg(const K &k, uint32_t n)
	{
		if (n == 0) return;
#ifndef B2J_HOSTSIM
		if (profiling) { K serial = k; serial.pdl = 0; launch_cfg<K, THREADS, MINB>(serial, n); return; }
		++launches;
		cudaLaunchConfig_t cfg = {};
		cfg.gridDim = dim3(grid_for(n, THREADS)); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
		cudaLaunchAttribute attr[1];
		attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		attr[0].val.programmaticStreamSerializationAllowed = 1;
		cfg.attrs = attr; cfg.numAttrs = 1;
		cudaLaunchKernelEx(&cfg, run_kernel_cfg<K, THREADS, MINB>, k, n);
#else
		launch(k, n);
#endif
	}
	template <class K, int THREADS, int MINB> void launch_cfg(const K &k, uint32_t n)
	{
		if (n == 0) return;
		++launches;
#ifndef B2J_HOSTSIM
		if (profiling) prof_begin(profile_category<K>());
		run_kernel_cfg<K, THREADS, MINB><<<grid_for(n, THREADS), THREADS, 0, stream>>>(k, n);
		if (profiling) prof_end();
#else
		for (uint32_t i = 0; i < n; ++i) k(i);
#endif
	}
	// device-resident count (no host sync): processes [*begin, min(*n_ptr, cap))
	template <class K> void launch_dev(const K &k, const uint32_t *n_ptr, const uint32_t *begin_ptr, uint32_t cap, uint32_t blocks_per_sm = 8)
	{
		if (cap == 0) return;
		++launches;
#ifndef B2J_HOSTSIM
		uint32_t g = grid_for(cap, 128);
		uint32_t gmax = (uint32_t)num_sms * blocks_per_sm; // (16 blocks per SM for every kernel measured 1% slower on the 4096 world batch)
		if (profiling) prof_begin(profile_category<K>());
		run_kernel_dev<K><<<g > gmax? gmax : g, 128, 0, stream>>>(k, n_ptr, begin_ptr, cap);
		if (profiling) prof_end();
#else
		uint32_t n = *n_ptr < cap? *n_ptr : cap;
		uint32_t begin = begin_ptr? *begin_ptr : 0;
		for (uint32_t i = 0; begin + i < n; ++i) k(i);
#endif
	}
	// device-resident count, one warp per item with an S scratch block in shared memory; uses at most num_slots warps
	template <class K> void launch_dev_lockstep(const K &k, const uint32_t *n_ptr, const uint32_t *begin_ptr, uint32_t cap)
	{
		if (cap == 0) return;
		++launches;
#ifndef B2J_HOSTSIM
		uint32_t g = grid_for(cap, 128);
		uint32_t gmax = (uint32_t)num_sms * 8; // (16 blocks per SM measured 1% slower on the 4096 world batch)
		if (profiling) prof_begin(profile_category<K>());
		run_kernel_dev_lockstep<K><<<g > gmax? gmax : g, 128, 0, stream>>>(k, n_ptr, begin_ptr, cap);
		if (profiling) prof_end();
#else
		uint32_t n = *n_ptr < cap? *n_ptr : cap;
		uint32_t begin = begin_ptr? *begin_ptr : 0;
		for (uint32_t i = begin; i < n; ++i) k.run(i - begin, true);
#endif
	}
	template <class K, class S> void launch_warp_smem(const K &k, const uint32_t *n_ptr, uint32_t cap, uint32_t num_slots, uint32_t warps_per_block = 4)
	{
		if (cap == 0) return;
		++launches;
#ifndef B2J_HOSTSIM
		// (per Runtime = per device and per launching thread: a function attribute is a per device setting, and the groups of a batch launch
		// from their own host threads)
		bool &configured = func_configured[(const void *)run_kernel_warp_smem<K, S>];
		if (!configured)
		{
			cudaFuncSetAttribute(run_kernel_warp_smem<K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * sizeof(S) < 227 * 1024? 8 * sizeof(S) : 4 * sizeof(S)));
			configured = true;
		}
		uint32_t g = num_slots / warps_per_block;
		if (g < 1) g = 1;
		uint32_t gn = (cap + warps_per_block - 1) / warps_per_block;
		if (gn < g) g = gn;
		if (profiling) prof_begin(profile_category<K>());
		run_kernel_warp_smem<K, S><<<g, 32 * warps_per_block, warps_per_block * sizeof(S), stream>>>(k, n_ptr, cap);
		if (profiling) prof_end();
#else
		(void)num_slots;
		static S scratch;
		uint32_t n = *n_ptr < cap? *n_ptr : cap;
		for (uint32_t i = 0; i < n; ++i) k.run(i, true, 0, scratch);
#endif
	}

	// KW = the warp cooperative form of the item body (the device runs it), KS = its serial statement (what the host simulation runs)
	template <class KW, class KS, class S, class H> void launch_warp_coop(const KW &kw, const KS &ks, const uint32_t *n_ptr, uint32_t cap, uint32_t num_slots, uint32_t warps_per_block = 4)
	{
		if (cap == 0) return;
		++launches;
#ifndef B2J_HOSTSIM
		(void)ks;
		static_assert(sizeof(S) % 16 == 0, "scratch blocks are packed back to back");
		size_t smem = (size_t)warps_per_block * (sizeof(S) + sizeof(H));
		bool &configured = func_configured[(const void *)run_kernel_warp_coop<KW, S, H>];
		if (!configured)
		{
			cudaFuncSetAttribute(run_kernel_warp_coop<KW, S, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
			configured = true;
		}
		uint32_t g = num_slots / warps_per_block;
		if (g < 1) g = 1;
		uint32_t gn = (cap + warps_per_block - 1) / warps_per_block;
		if (gn < g) g = gn;
		if (profiling) prof_begin(profile_category<KS>());
		run_kernel_warp_coop<KW, S, H><<<g, 32 * warps_per_block, smem, stream>>>(kw, n_ptr, cap);
		if (profiling) prof_end();
#else
		(void)num_slots; (void)kw; (void)warps_per_block;
		static S scratch;
		uint32_t n = *n_ptr < cap? *n_ptr : cap;
		for (uint32_t i = 0; i < n; ++i) ks.run(i, true, 0, scratch);
#endif
	}

	template <class K, class S> void launch_lane_local(const K &k, const uint32_t *n_ptr, uint32_t cap, uint32_t blocks_per_sm)
	{
		if (cap == 0) return;
		++launches;
#ifndef B2J_HOSTSIM
		uint32_t g = (uint32_t)num_sms * blocks_per_sm;
		uint32_t gn = (cap + 127) / 128;
		if (gn < g) g = gn;
		if (profiling) prof_begin(profile_category<K>());
		run_kernel_lane_local<K, S><<<g, 128, 0, stream>>>(k, n_ptr, cap);
		if (profiling) prof_end();
#else
		(void)blocks_per_sm;
		static S scratch;
		uint32_t n = *n_ptr < cap? *n_ptr : cap;
		for (uint32_t i = 0; i < n; ++i) k.run(i, true, 0, scratch);
#endif
	}

	// ---- persistent staging: one device buffer + one pinned host mirror, bump allocated per API call (no cudaMalloc per call)
	unsigned char *stage_dev = nullptr, *stage_host = nullptr;
	size_t stage_cap = 0, stage_used = 0;
	void stage_begin(size_t bytes)
	{
		bytes += 256 * 16;
		if (bytes > stage_cap)
		{
			sync();
			size_t ncap = bytes + bytes / 2;
#ifndef B2J_HOSTSIM
			if (stage_dev) cudaFree(stage_dev);
			if (stage_host) cudaFreeHost(stage_host);
			cudaMalloc((void **)&stage_dev, ncap);
			cudaMallocHost((void **)&stage_host, ncap);
#else
			free(stage_dev); free(stage_host);
			stage_dev = (unsigned char *)malloc(ncap);
			stage_host = (unsigned char *)malloc(ncap);
#endif
			stage_cap = ncap;
		}
		stage_used = 0;
	}
	// returns the device pointer; *host receives the pinned mirror of the same region
	template <class T> T *stage_alloc(size_t n, T **host)
	{
		size_t off = (stage_used + 255) & ~(size_t)255;
		stage_used = off + n * sizeof(T);
		*host = reinterpret_cast<T *>(stage_host + off);
		return reinterpret_cast<T *>(stage_dev + off);
	}
	void stage_to_device(size_t begin, size_t end)
	{
		if (end <= begin) return;
#ifndef B2J_HOSTSIM
		cudaMemcpyAsync(stage_dev + begin, stage_host + begin, end - begin, cudaMemcpyHostToDevice, stream);
#else
		memcpy(stage_dev + begin, stage_host + begin, end - begin);
#endif
	}
	// page locked host memory of the caller (cudaHostAlloc / cudaHostRegister / torch pin_memory)? Then copies go straight between it
	// and the device instead of through the pinned mirror (one host memcpy less per direction)
	bool is_pinned(const void *p) const
	{
#ifndef B2J_HOSTSIM
		if (p == nullptr) return false;
		cudaPointerAttributes a;
		if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
		return a.type == cudaMemoryTypeHost;
#else
		(void)p;
		return false;
#endif
	}
	void copy_to_device_async(void *dev, const void *host, size_t bytes)
	{
#ifndef B2J_HOSTSIM
		cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, stream);
#else
		memcpy(dev, host, bytes);
#endif
	}
	void copy_to_host_async(void *host, const void *dev, size_t bytes)
	{
#ifndef B2J_HOSTSIM
		cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, stream);
#else
		memcpy(host, dev, bytes);
#endif
	}
	void stage_to_host(size_t begin, size_t end)
	{
		if (end <= begin) return;
#ifndef B2J_HOSTSIM
		cudaMemcpyAsync(stage_host + begin, stage_dev + begin, end - begin, cudaMemcpyDeviceToHost, stream);
		cudaStreamSynchronize(stream);
#else
		memcpy(stage_host + begin, stage_dev + begin, end - begin);
#endif
	}

	void ensure_temp(size_t bytes)
	{
		if (bytes <= cub_temp_size) return;
#ifndef B2J_HOSTSIM
		if (cub_temp) { cudaStreamSynchronize(stream); cudaFree(cub_temp); }
		cudaMalloc(&cub_temp, bytes);
#else
		free(cub_temp);
		cub_temp = malloc(bytes);
#endif
		cub_temp_size = bytes;
	}

	// Sizes the CUB temporary storage once for the largest sorts / scans the world can issue: growing it on demand means a cudaFree +
	// cudaMalloc (device wide synchronisation, measured up to 0.5 s with > 100 GB allocated) whenever a count creeps up during a run.
	void reserve_temp(uint32_t max_sort64, uint32_t max_sort32, uint32_t max_scan)
	{
#ifndef B2J_HOSTSIM
		size_t need = 0, bytes = 0;
		cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)max_sort64, 0, 64, stream);
		need = bytes > need? bytes : need;
		cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)max_sort32, 0, 32, stream);
		need = bytes > need? bytes : need;
		cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)max_scan, stream);
		need = bytes > need? bytes : need;
		ensure_temp(need);
#else
		(void)max_sort64; (void)max_sort32; (void)max_scan;
#endif
	}

	// stable sort of (key, value) pairs, n known on the host; results in keys_out / vals_out
	template <class KeyT> void sort_pairs(const KeyT *keys_in, KeyT *keys_out, const uint32_t *vals_in, uint32_t *vals_out, uint32_t n, int end_bit = (int)sizeof(KeyT) * 8)
	{
		if (n == 0) return;
		launches += 1;
#ifndef B2J_HOSTSIM
		size_t bytes = 0;
		cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0, end_bit, stream);
		ensure_temp(bytes);
		cub::DeviceRadixSort::SortPairs(cub_temp, bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0, end_bit, stream);
#else
		(void)end_bit;
		std::vector<uint32_t> idx(n);
		std::iota(idx.begin(), idx.end(), 0u);
		std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return keys_in[a] < keys_in[b]; });
		for (uint32_t i = 0; i < n; ++i) { keys_out[i] = keys_in[idx[i]]; vals_out[i] = vals_in[idx[i]]; }
#endif
	}

	void exclusive_scan(const uint32_t *in, uint32_t *out, uint32_t n)
	{
		if (n == 0) return;
		launches += 1;
#ifndef B2J_HOSTSIM
		size_t bytes = 0;
		cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, stream);
		ensure_temp(bytes);
		cub::DeviceScan::ExclusiveSum(cub_temp, bytes, in, out, (int)n, stream);
#else
		uint32_t s = 0;
		for (uint32_t i = 0; i < n; ++i) { uint32_t v = in[i]; out[i] = s; s += v; }
#endif
	}
};

} // namespace b2j
