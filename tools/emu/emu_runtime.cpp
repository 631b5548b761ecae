// TEST INFRASTRUCTURE -- the CUDA-runtime side of the CPU emulation build (tools/emu): host memory as device memory, and
// streams that really are queues.
//
// Every asynchronous operation (kernel, cudaMemcpyAsync, event record / wait) is appended to its stream's queue and
// executed later, when the host synchronises (cudaStreamSynchronize, cudaEventSynchronize, cudaFree): the executor then
// runs whatever is runnable - an operation at the head of its stream whose awaited events have completed - until nothing
// is left.  Streams are visited newest first by default (B2R_EMU_ORDER=asc|desc|random:<seed>), i.e. work of a later
// frame is preferred over work of an earlier one whenever the product's events allow it: a missing dependency between
// frames in flight shows up as a wrong frame, not just as a theoretical race.
//
// Several contexts of one process stand for the ranks of a sharded group (B2R_TRACE_DEVICES=N; IPC handles carry plain
// addresses).  Their flag barriers work because a kernel may give up the processor: __nanosleep (the product calls it only
// inside the spin loop of peer_barrier_kernel) suspends the running kernel, the executor goes on with the other
// streams - including the peers' - and resumes it later.  Work that can never become runnable is reported as a deadlock.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <random>
#include <string>
#include <ucontext.h>
#include <vector>

namespace {

typedef int cudaError_t;
struct Event { unsigned flags; uint64_t enqueued = 0, completed = 0; };
struct KernelRun
{
	ucontext_t ctx;                 // where to resume (initially: the trampoline on `stack`)
	char* stack = nullptr;          // from a pool: most kernels run to completion at once and hand it straight back
	void (*fn)(void*); void* arg; void (*del)(void*);
	bool started = false, finished = false;
};
struct Op
{
	enum Type { KERNEL, COPY, MEMSET, RECORD, WAIT } type;
	KernelRun* kernel = nullptr;
	void* dst = nullptr; const void* src = nullptr; size_t bytes = 0; int value = 0;
	Event* event = nullptr; uint64_t seq = 0;
};
struct Stream { int id, dev; std::deque<Op> q; };

std::vector<Stream*> g_streams;
thread_local int t_device = 0;
ucontext_t g_executor;
KernelRun* g_running = nullptr;
int g_nextStream = 1;
constexpr size_t kKernelStack = 4u << 20;
std::vector<char*> g_stackPool;

void kernelTrampoline()
{
	KernelRun* k = g_running;
	k->fn(k->arg);
	k->finished = true;
	swapcontext(&k->ctx, &g_executor);
}
// runs / resumes the kernel at the head of a stream; true when it has finished
bool stepKernel(KernelRun* k)
{
	if (!k->started)
	{
		k->started = true;
		if (g_stackPool.empty()) k->stack = static_cast<char*>(malloc(kKernelStack)); else { k->stack = g_stackPool.back(); g_stackPool.pop_back(); }
		getcontext(&k->ctx);
		k->ctx.uc_stack.ss_sp = k->stack;
		k->ctx.uc_stack.ss_size = kKernelStack;
		k->ctx.uc_link = &g_executor;
		makecontext(&k->ctx, (void (*)())kernelTrampoline, 0);
	}
	g_running = k;
	swapcontext(&g_executor, &k->ctx);
	g_running = nullptr;
	if (k->finished && k->stack) { g_stackPool.push_back(k->stack); k->stack = nullptr; }
	return k->finished;
}

std::vector<Stream*> visitOrder()
{
	std::vector<Stream*> v = g_streams;
	const char* e = getenv("B2R_EMU_ORDER");
	const std::string mode = e ? e : "desc";
	if (mode == "desc") std::reverse(v.begin(), v.end());
	else if (mode.rfind("random", 0) == 0)
	{
		static std::mt19937 rng((unsigned)atoi(mode.c_str() + (mode.size() > 7 ? 7 : 6)) + 1u);
		std::shuffle(v.begin(), v.end(), rng);
	}
	return v;
}

// run everything that can run; returns when all queues are empty or nothing is runnable
void drain()
{
	if (g_running) { fprintf(stderr, "emu runtime: host synchronisation from inside a kernel\n"); abort(); }
	for (;;)
	{
		bool progress = false, pending = false;
		for (Stream* s : visitOrder())
		{
			// one operation per visit, so that the streams really interleave
			if (s->q.empty()) continue;
			pending = true;
			Op& op = s->q.front();
			bool done = false;
			switch (op.type)
			{
			case Op::WAIT: done = op.event->completed >= op.seq || getenv("B2R_EMU_IGNORE_WAITS"); break;      // (the knob shows that the waits matter)
			case Op::RECORD: op.event->completed = std::max(op.event->completed, op.seq); done = true; break;
			case Op::COPY: if (op.bytes) memmove(op.dst, op.src, op.bytes); done = true; break;
			case Op::MEMSET: if (op.bytes) memset(op.dst, op.value, op.bytes); done = true; break;
			case Op::KERNEL:
				done = stepKernel(op.kernel);
				if (done) { if (op.kernel->del) op.kernel->del(op.kernel->arg); delete op.kernel; }
				break;
			}
			if (done) { s->q.pop_front(); progress = true; }
		}
		if (!pending) return;
		if (!progress)
		{
			// only blocked waits and suspended kernels are left: give the suspended kernels a few more turns (a barrier
			// kernel polls a flag another suspended kernel may already have written), then stop
			static int idle = 0;
			if (++idle > 64) { idle = 0; return; }
		}
	}
}
void requireEmpty(Stream* s, const char* what)
{
	if (s && !s->q.empty())
	{
		fprintf(stderr, "emu runtime: deadlock - %s: stream %d (device %d) still has %zu operations that can never run\n", what, s->id, s->dev, s->q.size());
		static const char* names[] = { "kernel", "copy", "memset", "record", "wait" };
		for (Stream* t : g_streams)
			if (!t->q.empty())
			{
				const Op& op = t->q.front();
				fprintf(stderr, "  stream %d (device %d): %zu ops, head = %s", t->id, t->dev, t->q.size(), names[op.type]);
				if (op.type == Op::WAIT) fprintf(stderr, " event %p seq %llu (completed %llu)", (void*)op.event, (unsigned long long)op.seq, (unsigned long long)op.event->completed);
				if (op.type == Op::KERNEL) fprintf(stderr, " fn %p started %d", (void*)op.kernel->fn, (int)op.kernel->started);
				fprintf(stderr, "\n");
			}
		abort();
	}
}
Stream* legacy(int dev)
{
	for (Stream* s : g_streams) if (s->id == -1 - dev) return s;
	Stream* s = new Stream{ -1 - dev, dev, {} };
	g_streams.push_back(s);
	return s;
}
Stream* streamOf(void* s) { return s ? static_cast<Stream*>(s) : legacy(t_device); }

}

extern "C" {

// ---- the kernel side (tools/emu/emu_cuda.h)
void emu_enqueue_kernel(void* stream, void (*fn)(void*), void* arg, void (*del)(void*))
{
	Op op; op.type = Op::KERNEL;
	op.kernel = new KernelRun{};
	op.kernel->fn = fn; op.kernel->arg = arg; op.kernel->del = del;
	streamOf(stream)->q.push_back(op);
}
void emu_kernel_yield(void)
{
	KernelRun* k = g_running;
	if (!k) return;
	swapcontext(&k->ctx, &g_executor);
}
void b2r_trace_host_barrier(void) {}
void b2r_stub_noop(void) {}

// ---- device
cudaError_t cudaGetDeviceCount(int* n) { const char* e = getenv("B2R_TRACE_DEVICES"); *n = e ? atoi(e) : 1; return 0; }
cudaError_t cudaSetDevice(int d) { t_device = d; return 0; }
cudaError_t cudaDeviceGetAttribute(int* v, int attr, int) { *v = attr == 16 ? 148 : attr == 75 ? 10 : attr == 76 ? 0 : 1; return 0; }
cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { if (lo) *lo = 0; if (hi) *hi = -5; return 0; }
cudaError_t cudaFuncSetAttribute(const void*, int, int) { return 0; }
cudaError_t cudaGetLastError(void) { return 0; }
const char* cudaGetErrorString(cudaError_t e) { return e ? "emu runtime: unsupported" : "no error"; }

// ---- memory
cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
cudaError_t cudaFree(void* p) { drain(); (void)p; return 0; }          // synchronises the device; the block is kept (no address reuse)
cudaError_t cudaMallocHost(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
cudaError_t cudaFreeHost(void*) { return 0; }
// mapped host memory: "device" memory is host memory here, so the device view of a registered range is the range itself
cudaError_t cudaHostRegister(void*, size_t, unsigned) { return 0; }
cudaError_t cudaHostUnregister(void*) { return 0; }
cudaError_t cudaHostGetDevicePointer(void** dev, void* host, unsigned) { *dev = host; return 0; }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, int, void* stream)
{
	Op op; op.type = Op::COPY; op.dst = dst; op.src = src; op.bytes = n;
	streamOf(stream)->q.push_back(op);
	return 0;
}
// blocking calls on the legacy stream: ordered behind the legacy stream only - NOT behind the context's non-blocking streams
cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, int)
{
	Stream* l = legacy(t_device);
	Op op; op.type = Op::COPY; op.dst = dst; op.src = src; op.bytes = n;
	l->q.push_back(op);
	while (!l->q.empty())
	{
		Op& o = l->q.front();
		if (o.type == Op::KERNEL) { if (!stepKernel(o.kernel)) { fprintf(stderr, "emu runtime: a legacy-stream kernel waits for other work\n"); abort(); } delete o.kernel; }
		else if (o.type == Op::COPY) { if (o.bytes) memmove(o.dst, o.src, o.bytes); }
		else if (o.type == Op::MEMSET) { if (o.bytes) memset(o.dst, o.value, o.bytes); }
		l->q.pop_front();
	}
	return 0;
}
cudaError_t cudaMemset(void* p, int v, size_t n)
{
	Op op; op.type = Op::MEMSET; op.dst = p; op.value = v; op.bytes = n;      // asynchronous with respect to the host, on the legacy stream
	legacy(t_device)->q.push_back(op);
	return 0;
}
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, const void*, int, size_t) { *n = 0; return 1; }
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(int* n, const void*, int, size_t, unsigned) { *n = 0; return 1; }      // "unsupported": the product falls back to its launch bounds
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, void* stream)
{
	Op op; op.type = Op::MEMSET; op.dst = p; op.value = v; op.bytes = n;
	streamOf(stream)->q.push_back(op);
	return 0;
}

// ---- streams and events
cudaError_t cudaStreamCreateWithFlags(void** s, unsigned) { Stream* st = new Stream{ g_nextStream++, t_device, {} }; g_streams.push_back(st); *s = st; return 0; }
cudaError_t cudaStreamCreateWithPriority(void** s, unsigned f, int) { return cudaStreamCreateWithFlags(s, f); }
cudaError_t cudaStreamDestroy(void*) { return 0; }
cudaError_t cudaStreamSynchronize(void* s) { drain(); requireEmpty(streamOf(s), "cudaStreamSynchronize"); return 0; }
cudaError_t cudaEventCreateWithFlags(void** e, unsigned flags) { *e = new Event{ flags }; return 0; }
cudaError_t cudaEventCreate(void** e) { return cudaEventCreateWithFlags(e, 0); }
cudaError_t cudaEventDestroy(void*) { return 0; }
cudaError_t cudaEventRecord(void* e, void* s)
{
	Event* ev = static_cast<Event*>(e);
	Op op; op.type = Op::RECORD; op.event = ev; op.seq = ++ev->enqueued;
	streamOf(s)->q.push_back(op);
	return 0;
}
cudaError_t cudaStreamWaitEvent(void* s, void* e, unsigned)
{
	Event* ev = static_cast<Event*>(e);
	Op op; op.type = Op::WAIT; op.event = ev; op.seq = ev->enqueued;          // the most recent record call at this point
	streamOf(s)->q.push_back(op);
	return 0;
}
cudaError_t cudaEventSynchronize(void* e)
{
	drain();
	Event* ev = static_cast<Event*>(e);
	if (ev->completed < ev->enqueued) { fprintf(stderr, "emu runtime: deadlock - cudaEventSynchronize on an event whose record can never run\n"); abort(); }
	return 0;
}
cudaError_t cudaEventElapsedTime(float* ms, void* a, void* b)
{
	const Event* ea = static_cast<Event*>(a); const Event* eb = static_cast<Event*>(b);
	if ((ea->flags & 2u) || (eb->flags & 2u) || !ea->enqueued || !eb->enqueued) return 400;
	*ms = 0.01f;
	return 0;
}

// ---- IPC between the contexts of this process
struct IpcHandle { char reserved[64]; };
cudaError_t cudaIpcGetMemHandle(IpcHandle* h, void* p) { if (!h || !p) return 1; memset(h, 0, sizeof *h); memcpy(h->reserved, &p, sizeof p); return 0; }
cudaError_t cudaIpcOpenMemHandle(void** p, IpcHandle h, unsigned) { memcpy(p, h.reserved, sizeof *p); return *p ? 0 : 1; }
cudaError_t cudaIpcCloseMemHandle(void*) { return 0; }

}
