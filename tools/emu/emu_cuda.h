// TEST INFRASTRUCTURE -- a small CUDA execution model for the host compiler, so that the product's kernels can be RUN on
// a machine without a GPU (tools/emu/build.sh -> build/emu/libb2r_emu.so, selected in tests with B2R_LIBRARY).
//
// Never part of the product: libb2r.so has no CPU path and fails without a CUDA device.  This header exists to execute
// the very same kernel sources (csrc/*.cuh, unmodified; b2r.cu with its launch syntax rewritten by tools/emu/preprocess.py)
// functionally, e.g. to compare the frames of scheduler modes that were never run on a GPU with the default mode and
// with the oracle.  Memory, streams and events come from tools/emu/emu_runtime.cpp (host memory; streams are real queues,
// drained when the host synchronises, newest stream first); a kernel executes block after block in index order.
//
// Execution model: every CUDA thread of a block is a fiber (ucontext) on ONE host thread.  A fiber runs until it reaches
// __syncthreads or a warp collective; the scheduler resumes the others, and resolves a barrier / collective when all
// participants have arrived (all live threads of the block; for *_sync(mask, ...) the lanes named in the mask).  That is
// a legal schedule under CUDA's independent thread scheduling; code that relied on lock-step execution without
// synchronisation would misbehave here, as it may on hardware.  A block in which nobody can make progress is reported
// as a deadlock.  Arithmetic: the host's IEEE single precision (compile with -ffp-contract=off; fmaf is a real FMA),
// __fdiv_rn / __frcp_rn / __fsqrt_rn are the correctly rounded host operations.
#pragma once
#define __DEVICE_LAUNCH_PARAMETERS_H__      /* the built-in variables are defined below */
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <tuple>
#include <type_traits>
#include <ucontext.h>
#include <utility>
#include <vector>

#undef __global__
#undef __device__
#undef __host__
#undef __forceinline__
#undef __shared__
#undef __constant__
#undef __launch_bounds__
#undef __align__
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __shared__ static             /* one block runs at a time: block-shared = one static instance */
#define __constant__ static
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))
#define EMU_ASM(...) do { } while (0)  /* inline PTX (griddepcontrol, prefetch, globaltimer): no meaning here */

extern "C" void emu_enqueue_kernel(void* stream, void (*fn)(void*), void* arg, void (*del)(void*));
extern "C" void emu_kernel_yield(void);
inline uint3 threadIdx{ 0, 0, 0 }, blockIdx{ 0, 0, 0 };
inline dim3 blockDim{ 1, 1, 1 }, gridDim{ 1, 1, 1 };

namespace emu {

enum Kind { NONE, SYNCTHREADS, SYNCTHREADS_OR, SYNCWARP, SHFL, SHFL_UP, SHFL_DOWN, SHFL_XOR, BALLOT, ANY, ALL, RED_MIN_U, RED_MAX_U, RED_MIN_S, RED_MAX_S, RED_ADD, MATCH_ANY };
struct Fiber
{
	ucontext_t ctx;
	uint3 tid;
	int linear = 0;
	bool done = false, blocked = false;
	Kind kind = NONE;
	unsigned mask = 0;
	unsigned long long value = 0, result = 0;
	int arg = 0;
};
struct Engine
{
	std::vector<Fiber> fibers;
	std::vector<char> stacks;
	ucontext_t main;
	Fiber* cur = nullptr;
	uint3 bIdx{ 0, 0, 0 };
	dim3 bDim{ 1, 1, 1 }, gDim{ 1, 1, 1 };
	std::vector<char> dynamicSmem;
	void (*body)(void*) = nullptr;
	void* bodyArg = nullptr;
	const char* kernelName = "?";
};
// one Engine per kernel in execution: a kernel that has given up the processor (__nanosleep, see emu_runtime.cpp) keeps
// its fibers while other kernels run
inline Engine*& current() { static Engine* e = nullptr; return e; }
inline Engine& E() { return *current(); }
inline std::vector<Engine*>& pool() { static std::vector<Engine*> p; return p; }
constexpr size_t kStack = 256 * 1024;

inline void yieldToScheduler() { Engine& e = E(); swapcontext(&e.cur->ctx, &e.main); }
inline unsigned long long collective(Kind k, unsigned mask, unsigned long long v, int arg = 0)
{
	Fiber* f = E().cur;
	f->kind = k; f->mask = mask; f->value = v; f->arg = arg; f->blocked = true;
	yieldToScheduler();
	return f->result;
}
inline void trampoline()
{
	Engine& e = E();
	e.body(e.bodyArg);
	e.cur->done = true;
	yieldToScheduler();
}

inline bool resolveWarp(Engine& e, int warp, int nThreads)
{
	bool any = false;
	const int base = warp * 32, n = std::min(32, nThreads - base);
	for (int l0 = 0; l0 < n; ++l0)
	{
		Fiber& a = e.fibers[base + l0];
		if (!a.blocked || a.kind <= SYNCTHREADS_OR) continue;
		// participants: the lanes named in the mask that exist and have not exited
		unsigned want = 0, have = 0;
		for (int l = 0; l < n; ++l)
			if ((a.mask >> l) & 1u)
			{
				Fiber& b = e.fibers[base + l];
				if (b.done) continue;
				want |= 1u << l;
				if (b.blocked && b.kind == a.kind && b.mask == a.mask) have |= 1u << l;
			}
		if (!(want >> l0 & 1u)) { fprintf(stderr, "emu: %s: lane %d calls a warp collective whose mask %#x does not name it\n", e.kernelName, l0, a.mask); abort(); }
		if (have != want) continue;
		auto val = [&](int l) -> unsigned long long { return e.fibers[base + l].value; };
		unsigned ballot = 0; unsigned long long acc = 0; bool first = true;
		for (int l = 0; l < n; ++l)
			if ((want >> l) & 1u)
			{
				const unsigned long long v = val(l);
				if (v) ballot |= 1u << l;
				switch (a.kind)
				{
				case RED_MIN_U: acc = first ? v : std::min(acc, v); break;
				case RED_MAX_U: acc = first ? v : std::max(acc, v); break;
				case RED_MIN_S: acc = first ? v : (unsigned long long)std::min((long long)acc, (long long)v); break;
				case RED_MAX_S: acc = first ? v : (unsigned long long)std::max((long long)acc, (long long)v); break;
				case RED_ADD: acc += v; break;
				default: break;
				}
				first = false;
			}
		for (int l = 0; l < n; ++l)
			if ((want >> l) & 1u)
			{
				Fiber& b = e.fibers[base + l];
				unsigned long long r = 0;
				int src;
				switch (a.kind)
				{
				case SYNCWARP: break;
				case SHFL: src = b.arg & 31; r = (src < n && ((want >> src) & 1u)) ? val(src) : b.value; break;
				case SHFL_UP: src = l - b.arg; r = (src >= 0 && ((want >> src) & 1u)) ? val(src) : b.value; break;
				case SHFL_DOWN: src = l + b.arg; r = (src < n && src < 32 && ((want >> src) & 1u)) ? val(src) : b.value; break;
				case SHFL_XOR: src = l ^ b.arg; r = (src < n && ((want >> src) & 1u)) ? val(src) : b.value; break;
				case BALLOT: r = ballot; break;
				case ANY: r = ballot != 0; break;
				case ALL: r = ballot == want; break;
				case RED_MIN_U: case RED_MAX_U: case RED_MIN_S: case RED_MAX_S: case RED_ADD: r = acc; break;
				case MATCH_ANY: { unsigned m = 0; for (int k = 0; k < n; ++k) if (((want >> k) & 1u) && val(k) == b.value) m |= 1u << k; r = m; break; }
				default: break;
				}
				b.result = r;
			}
		for (int l = 0; l < n; ++l)
			if ((want >> l) & 1u) { e.fibers[base + l].blocked = false; e.fibers[base + l].kind = NONE; }
		any = true;
	}
	return any;
}

inline void runBlock(Engine& e)
{
	const int n = (int)(e.bDim.x * e.bDim.y * e.bDim.z);
	if ((int)e.fibers.size() < n) e.fibers.resize(n);
	if (e.stacks.size() < (size_t)n * kStack) e.stacks.resize((size_t)n * kStack);
	for (int i = 0; i < n; ++i)
	{
		Fiber& f = e.fibers[i];
		f.linear = i; f.done = false; f.blocked = false; f.kind = NONE;
		f.tid.x = i % e.bDim.x; f.tid.y = (i / e.bDim.x) % e.bDim.y; f.tid.z = i / (e.bDim.x * e.bDim.y);
		getcontext(&f.ctx);
		f.ctx.uc_stack.ss_sp = e.stacks.data() + (size_t)i * kStack;
		f.ctx.uc_stack.ss_size = kStack;
		f.ctx.uc_link = &e.main;
		makecontext(&f.ctx, (void (*)())trampoline, 0);
	}
	int alive = n;
	while (alive)
	{
		bool progressed = false;
		for (int i = 0; i < n; ++i)
		{
			Fiber& f = e.fibers[i];
			if (f.done || f.blocked) continue;
			e.cur = &f;
			::threadIdx = f.tid;
			swapcontext(&e.main, &f.ctx);
			progressed = true;
			if (f.done) --alive;
		}
		// block-wide barrier: all live threads have arrived
		int waiting = 0; unsigned long long orv = 0; Kind bk = NONE;
		for (int i = 0; i < n; ++i)
		{
			Fiber& f = e.fibers[i];
			if (!f.done && f.blocked && (f.kind == SYNCTHREADS || f.kind == SYNCTHREADS_OR)) { ++waiting; orv |= f.value; bk = f.kind; }
		}
		if (alive && waiting == alive)
		{
			for (int i = 0; i < n; ++i)
				if (!e.fibers[i].done) { e.fibers[i].result = (bk == SYNCTHREADS_OR) ? (orv != 0) : 0; e.fibers[i].blocked = false; e.fibers[i].kind = NONE; }
			progressed = true;
		}
		for (int w = 0; w * 32 < n; ++w) progressed |= resolveWarp(e, w, n);
		if (!progressed && alive)
		{
			fprintf(stderr, "emu: deadlock in %s block (%u,%u,%u): %d threads alive, %d at __syncthreads\n", e.kernelName, e.bIdx.x, e.bIdx.y, e.bIdx.z, alive, waiting);
			abort();
		}
	}
}

template <typename F>
inline void runGrid(const char* name, dim3 grid, dim3 block, size_t dynSmem, F&& f)
{
	Engine* prev = current();
	Engine* ep;
	if (pool().empty()) ep = new Engine; else { ep = pool().back(); pool().pop_back(); }
	current() = ep;
	Engine& e = *ep;
	e.kernelName = name; e.gDim = grid; e.bDim = block;
	e.dynamicSmem.assign(dynSmem + 16, 0);
	auto* fp = &f;
	e.body = [](void* p) { (*static_cast<typename std::remove_reference<F>::type*>(p))(); };
	e.bodyArg = (void*)fp;
	for (unsigned z = 0; z < grid.z; ++z)
		for (unsigned y = 0; y < grid.y; ++y)
			for (unsigned x = 0; x < grid.x; ++x)
			{
				e.bIdx = uint3{ x, y, z };
				::blockIdx = e.bIdx; ::blockDim = block; ::gridDim = grid;
				runBlock(e);
			}
	pool().push_back(ep);
	current() = prev;
}
inline void* dynamicSmemPtr() { return E().dynamicSmem.data(); }
template <typename K> inline cudaError_t funcSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }

// kernel<<<grid, block[, smem, stream]>>>(args...)  (rewritten by tools/emu/preprocess.py) and cudaLaunchKernelEx: the launch
// is QUEUED on its stream (emu_runtime.cpp) with a copy of its arguments and runs when the executor gets to it
template <typename Tuple, typename K>
struct Closure { const char* name; K kernel; dim3 grid, block; size_t smem; Tuple params; };
template <typename... P, typename... A>
inline void launch(const char* name, void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args)
{
	using T = std::tuple<typename std::decay<P>::type...>;
	using C = Closure<T, void (*)(P...)>;
	C* c = new C{ name, kernel, grid, block, smem, T(std::forward<A>(args)...) };
	emu_enqueue_kernel((void*)stream,
		[](void* p) { C* c = static_cast<C*>(p); runGrid(c->name, c->grid, c->block, c->smem, [c] { std::apply(c->kernel, c->params); }); },
		c, [](void* p) { delete static_cast<C*>(p); });
}
template <typename... P, typename... A>
inline cudaError_t launchEx(const cudaLaunchConfig_t* cfg, void (*kernel)(P...), A&&... args)
{
	launch("kernel", kernel, cfg->gridDim, cfg->blockDim, cfg->dynamicSmemBytes, cfg->stream, std::forward<A>(args)...);
	return cudaSuccess;
}

}

// ---- built-in variables: plain globals (one host thread), refreshed by the scheduler before it resumes a fiber
// ---- synchronisation and warp collectives
inline void __syncthreads() { emu::collective(emu::SYNCTHREADS, 0, 0); }
inline int __syncthreads_or(int p) { return (int)emu::collective(emu::SYNCTHREADS_OR, 0, p != 0); }
inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) { emu::collective(emu::SYNCWARP, mask, 0); }
template <typename T> inline unsigned long long emuBits(T v) { unsigned long long b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> inline T emuFrom(unsigned long long b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
template <typename T> inline T __shfl_sync(unsigned m, T v, int src, int = 32) { return emuFrom<T>(emu::collective(emu::SHFL, m, emuBits(v), src)); }
template <typename T> inline T __shfl_up_sync(unsigned m, T v, unsigned d, int = 32) { return emuFrom<T>(emu::collective(emu::SHFL_UP, m, emuBits(v), (int)d)); }
template <typename T> inline T __shfl_down_sync(unsigned m, T v, unsigned d, int = 32) { return emuFrom<T>(emu::collective(emu::SHFL_DOWN, m, emuBits(v), (int)d)); }
template <typename T> inline T __shfl_xor_sync(unsigned m, T v, int d, int = 32) { return emuFrom<T>(emu::collective(emu::SHFL_XOR, m, emuBits(v), d)); }
inline unsigned __ballot_sync(unsigned m, int p) { return (unsigned)emu::collective(emu::BALLOT, m, p != 0); }
inline int __any_sync(unsigned m, int p) { return (int)emu::collective(emu::ANY, m, p != 0); }
inline int __all_sync(unsigned m, int p) { return (int)emu::collective(emu::ALL, m, p != 0); }
inline unsigned __reduce_min_sync(unsigned m, unsigned v) { return (unsigned)emu::collective(emu::RED_MIN_U, m, v); }
inline unsigned __reduce_max_sync(unsigned m, unsigned v) { return (unsigned)emu::collective(emu::RED_MAX_U, m, v); }
inline int __reduce_min_sync(unsigned m, int v) { return (int)(long long)emu::collective(emu::RED_MIN_S, m, (unsigned long long)(long long)v); }
inline int __reduce_max_sync(unsigned m, int v) { return (int)(long long)emu::collective(emu::RED_MAX_S, m, (unsigned long long)(long long)v); }
inline unsigned __reduce_add_sync(unsigned m, unsigned v) { return (unsigned)emu::collective(emu::RED_ADD, m, v); }
template <typename T> inline unsigned __match_any_sync(unsigned m, T v) { return (unsigned)emu::collective(emu::MATCH_ANY, m, emuBits(v)); }
inline void __threadfence() {}
inline void __threadfence_system() {}
inline void __threadfence_block() {}
// the product sleeps only inside the spin loop of its cross-GPU flag barrier: let the executor run the other streams (the
// peers' among them) and come back
inline void __nanosleep(unsigned)
{
	emu::Engine* e = emu::current();
	emu::Fiber* f = e->cur;
	emu_kernel_yield();
	emu::current() = e;
	::threadIdx = f->tid; ::blockIdx = e->bIdx; ::blockDim = e->bDim; ::gridDim = e->gDim;
}
inline unsigned long long emuGlobalTimerNs() { static unsigned long long t = 0; return t += 1000; }     // never reaches the 4 s time-out

// ---- atomics (one host thread: plain read-modify-write)
template <typename T> inline T atomicAdd(T* p, T v) { const T o = *p; *p = o + v; return o; }
template <typename T> inline T atomicMin(T* p, T v) { const T o = *p; *p = std::min(o, v); return o; }
template <typename T> inline T atomicMax(T* p, T v) { const T o = *p; *p = std::max(o, v); return o; }
template <typename T> inline T atomicCAS(T* p, T cmp, T v) { const T o = *p; if (o == cmp) *p = v; return o; }
template <typename T> inline T atomicExch(T* p, T v) { const T o = *p; *p = v; return o; }
template <typename T> inline T atomicOr(T* p, T v) { const T o = *p; *p = o | v; return o; }

// ---- loads, bit operations, conversions, IEEE arithmetic
template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __ldcs(const T* p) { return *p; }
template <typename T> inline void __stcs(T* p, T v) { *p = v; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i); return r; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __frcp_rn(float a) { return 1.0f / a; }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline unsigned __float_as_uint(float f) { return emuFrom<unsigned>(emuBits(f)); }
inline int __float_as_int(float f) { return emuFrom<int>(emuBits(f)); }
inline float __uint_as_float(unsigned u) { return emuFrom<float>(emuBits(u)); }
inline float __int_as_float(int u) { return emuFrom<float>(emuBits(u)); }
// cvt.*.s32.f32 / u32.f32: saturating, NaN -> 0
inline int __float2int_rz(float f) { if (f != f) return 0; if (f >= 2147483648.f) return INT32_MAX; if (f <= -2147483648.f) return INT32_MIN; return (int)f; }
inline int __float2int_rd(float f) { return __float2int_rz(floorf(f)); }
inline int __float2int_ru(float f) { return __float2int_rz(ceilf(f)); }
inline int __float2int_rn(float f) { return __float2int_rz(nearbyintf(f)); }
inline unsigned __float2uint_rz(float f) { if (!(f > 0.f)) return 0u; if (f >= 4294967296.f) return UINT32_MAX; return (unsigned)f; }
inline float __int2float_rn(int v) { return (float)v; }
inline float __uint2float_rn(unsigned v) { return (float)v; }

// min / max over mixed integer types, as CUDA's overload set allows
template <typename A, typename B, typename = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>::type>
inline typename std::common_type<A, B>::type min(A a, B b) { using C = typename std::common_type<A, B>::type; return (C)a < (C)b ? (C)a : (C)b; }
template <typename A, typename B, typename = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>::type>
inline typename std::common_type<A, B>::type max(A a, B b) { using C = typename std::common_type<A, B>::type; return (C)a > (C)b ? (C)a : (C)b; }
