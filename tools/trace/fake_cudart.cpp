// TEST INFRASTRUCTURE -- a recording stand-in for the CUDA runtime, for machines without a GPU.
//
// tools/trace/build.sh links the product's own host code (csrc/b2r.cu + the veneer, compiled with `nvcc -cudart none`)
// against this library instead of libcudart.  "Device" memory is host memory, copies are memcpy, kernels do not run;
// every stream / event / copy / launch the host code issues is appended as one JSON line to $B2R_TRACE_FILE, kernel
// launches with their argument bytes (parameter sizes come from the cubin: tools/trace/kernel_params.py writes the
// table named by $B2R_TRACE_PARAMS).  tools/trace/stream_order.py replays the log and checks that every pair of
// conflicting device accesses is ordered by stream order, events or host synchronisation - a race detector for the
// frame scheduler (enqueueFrame / enqueueFrameRing) that needs no GPU.  Never part of the product.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {

typedef int cudaError_t;
struct dim3 { unsigned x, y, z; };
struct Stream { int id; int dev; };
struct Event { int id; unsigned flags; bool recorded; };

std::mutex g_mu;
FILE* g_log = nullptr;
thread_local int t_device = 0;        // cudaSetDevice: which context ("rank") is issuing - several can share this process
int g_nextStream = 1, g_nextEvent = 1;
std::map<const void*, std::string> g_kernelName;                 // host stub address -> mangled device name
std::map<std::string, std::vector<std::pair<int, int>>> g_params; // mangled name -> (offset, size) by ordinal
bool g_paramsLoaded = false;

struct CallConfig { dim3 grid, block; size_t shmem; void* stream; };
thread_local std::vector<CallConfig> t_config;

FILE* log()
{
	if (!g_log)
	{
		const char* p = getenv("B2R_TRACE_FILE");
		g_log = p ? fopen(p, "a") : nullptr;
	}
	return g_log;
}
void loadParams()
{
	if (g_paramsLoaded) return;
	g_paramsLoaded = true;
	const char* p = getenv("B2R_TRACE_PARAMS");
	FILE* f = p ? fopen(p, "r") : nullptr;
	if (!f) return;
	char name[1024]; int n;
	while (fscanf(f, "%1023s %d", name, &n) == 2)
	{
		std::vector<std::pair<int, int>> v(n);
		for (int i = 0; i < n; ++i) if (fscanf(f, "%d %d", &v[i].first, &v[i].second) != 2) { fclose(f); return; }
		g_params[name] = v;
	}
	fclose(f);
}
int streamId(void* s) { return s ? static_cast<Stream*>(s)->id : 0; }
// every line ends with the issuing device: {"op":...,"dev":d}
#define LOG(fmt, ...) do { std::lock_guard<std::mutex> lk_(g_mu); if (FILE* f_ = log()) { fprintf(f_, fmt ",\"dev\":%d}\n", ##__VA_ARGS__, t_device); fflush(f_); } } while (0)

void logLaunch(const void* func, dim3 grid, dim3 block, void** args, void* stream, int pdl)
{
	loadParams();
	std::lock_guard<std::mutex> lk(g_mu);
	FILE* f = log();
	if (!f) return;
	auto it = g_kernelName.find(func);
	const std::string name = it == g_kernelName.end() ? "?" : it->second;
	fprintf(f, "{\"op\":\"launch\",\"kernel\":\"%s\",\"stream\":%d,\"grid\":[%u,%u,%u],\"block\":[%u,%u,%u],\"pdl\":%d,\"args\":[", name.c_str(), streamId(stream),
		grid.x, grid.y, grid.z, block.x, block.y, block.z, pdl);
	// the one kernel whose RESULT steers the host code: prim_vertex_range_kernel(triRefs, primLo, primHi, out) - the range of
	// vertices a rank's primitive slice refers to (sliced vertex stage).  Stand-in result: "all of them", so that the
	// scheduler goes on to launch the vertex stage for its chunks
	if (name.find("prim_vertex_range_kernel") != std::string::npos && args)
	{
		uint32_t* out = *static_cast<uint32_t**>(args[3]);
		out[0] = 0u; out[1] = 0xFFFFFFF0u;
	}
	auto pit = g_params.find(name);
	if (pit != g_params.end() && args)
		for (size_t i = 0; i < pit->second.size(); ++i)
		{
			fprintf(f, "%s\"", i ? "," : "");
			const unsigned char* a = static_cast<const unsigned char*>(args[i]);
			for (int b = 0; b < pit->second[i].second; ++b) fprintf(f, "%02x", a[b]);
			fputc('"', f);
		}
	fprintf(f, "],\"dev\":%d}\n", t_device);
	fflush(f);
}

}

extern "C" {

// ---- registration (what nvcc's host stubs call at load time)
void** __cudaRegisterFatBinary(void*) { static void* h; return &h; }
void __cudaRegisterFatBinaryEnd(void**) {}
void __cudaUnregisterFatBinary(void**) {}
void __cudaRegisterFunction(void**, const char* hostFun, char*, const char* deviceName, int, void*, void*, void*, void*, int*)
{
	std::lock_guard<std::mutex> lk(g_mu);
	g_kernelName[hostFun] = deviceName;
}
unsigned __cudaPushCallConfiguration(dim3 grid, dim3 block, size_t shmem, void* stream)
{
	t_config.push_back(CallConfig{ grid, block, shmem, stream });
	return 0;
}
cudaError_t __cudaPopCallConfiguration(dim3* grid, dim3* block, size_t* shmem, void* stream)
{
	if (t_config.empty()) return 1;
	const CallConfig c = t_config.back(); t_config.pop_back();
	*grid = c.grid; *block = c.block; *shmem = c.shmem; *static_cast<void**>(stream) = c.stream;
	return 0;
}

// ---- device
cudaError_t cudaGetDeviceCount(int* n)
{
	const char* e = getenv("B2R_TRACE_DEVICES");       // several contexts of one process can stand for the ranks of a sharded group
	*n = e ? atoi(e) : 1;
	return 0;
}
cudaError_t cudaSetDevice(int d) { t_device = d; return 0; }
cudaError_t cudaDeviceGetAttribute(int* v, int attr, int)
{
	switch (attr)
	{
	case 16: *v = 148; break;          // cudaDevAttrMultiProcessorCount
	case 75: *v = 10; break;           // cudaDevAttrComputeCapabilityMajor
	case 76: *v = 0; break;            // cudaDevAttrComputeCapabilityMinor
	default: *v = 1; break;
	}
	return 0;
}
cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { if (lo) *lo = 0; if (hi) *hi = -5; return 0; }
cudaError_t cudaFuncSetAttribute(const void*, int, int) { return 0; }
cudaError_t cudaGetLastError(void) { return 0; }
const char* cudaGetErrorString(cudaError_t e) { return e ? "fake runtime: unsupported" : "no error"; }

// ---- memory ("device" memory is host memory)
cudaError_t cudaMalloc(void** p, size_t n)
{
	*p = calloc(n ? n : 1, 1);
	LOG("{\"op\":\"malloc\",\"ptr\":%llu,\"size\":%zu", (unsigned long long)(uintptr_t)*p, n);
	return *p ? 0 : 2;
}
cudaError_t cudaFree(void* p)
{
	if (p) LOG("{\"op\":\"free\",\"ptr\":%llu", (unsigned long long)(uintptr_t)p);      // cudaFree synchronises the device
	// the block is not returned to the heap: the log identifies allocations by address, and a recycled address would
	// alias two different buffers
	return 0;
}
cudaError_t cudaMallocHost(void** p, size_t n)
{
	*p = calloc(n ? n : 1, 1);
	LOG("{\"op\":\"malloc_host\",\"ptr\":%llu,\"size\":%zu", (unsigned long long)(uintptr_t)*p, n);
	return *p ? 0 : 2;
}
cudaError_t cudaFreeHost(void* p) { if (p) LOG("{\"op\":\"free_host\",\"ptr\":%llu", (unsigned long long)(uintptr_t)p); return 0; }
// a registered (mapped) host range is logged like a pinned allocation; its device view is the range itself
cudaError_t cudaHostRegister(void* p, size_t n, unsigned) { LOG("{\"op\":\"malloc_host\",\"ptr\":%llu,\"size\":%zu", (unsigned long long)(uintptr_t)p, n); return 0; }
cudaError_t cudaHostUnregister(void* p) { if (p) LOG("{\"op\":\"free_host\",\"ptr\":%llu", (unsigned long long)(uintptr_t)p); return 0; }
cudaError_t cudaHostGetDevicePointer(void** dev, void* host, unsigned) { *dev = host; return 0; }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, int kind, void* stream)
{
	if (n) memmove(dst, src, n);
	LOG("{\"op\":\"memcpy\",\"dst\":%llu,\"src\":%llu,\"bytes\":%zu,\"kind\":%d,\"stream\":%d,\"async\":1", (unsigned long long)(uintptr_t)dst, (unsigned long long)(uintptr_t)src, n, kind, streamId(stream));
	return 0;
}
cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, int kind)
{
	if (n) memmove(dst, src, n);
	LOG("{\"op\":\"memcpy\",\"dst\":%llu,\"src\":%llu,\"bytes\":%zu,\"kind\":%d,\"stream\":0,\"async\":0", (unsigned long long)(uintptr_t)dst, (unsigned long long)(uintptr_t)src, n, kind);
	return 0;
}
cudaError_t cudaMemset(void* p, int v, size_t n)
{
	if (n) memset(p, v, n);
	LOG("{\"op\":\"memset\",\"dst\":%llu,\"bytes\":%zu,\"stream\":0", (unsigned long long)(uintptr_t)p, n);
	return 0;
}
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, const void*, int, size_t) { *n = 0; return 1; }
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(int* n, const void*, int, size_t, unsigned) { *n = 0; return 1; }
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, void* stream)
{
	if (n) memset(p, v, n);
	LOG("{\"op\":\"memset\",\"dst\":%llu,\"bytes\":%zu,\"stream\":%d", (unsigned long long)(uintptr_t)p, n, streamId(stream));
	return 0;
}

// ---- streams and events
cudaError_t cudaStreamCreateWithFlags(void** s, unsigned flags)
{
	Stream* st = new Stream{ g_nextStream++, t_device };
	*s = st;
	LOG("{\"op\":\"stream_create\",\"stream\":%d,\"flags\":%u", st->id, flags);
	return 0;
}
cudaError_t cudaStreamCreateWithPriority(void** s, unsigned flags, int) { return cudaStreamCreateWithFlags(s, flags); }
cudaError_t cudaStreamDestroy(void*) { return 0; }
cudaError_t cudaStreamSynchronize(void* s) { LOG("{\"op\":\"stream_sync\",\"stream\":%d", streamId(s)); return 0; }
cudaError_t cudaEventCreateWithFlags(void** e, unsigned flags) { *e = new Event{ g_nextEvent++, flags, false }; return 0; }
cudaError_t cudaEventCreate(void** e) { return cudaEventCreateWithFlags(e, 0); }
cudaError_t cudaEventDestroy(void*) { return 0; }
cudaError_t cudaEventRecord(void* e, void* s) { static_cast<Event*>(e)->recorded = true; LOG("{\"op\":\"event_record\",\"event\":%d,\"stream\":%d", static_cast<Event*>(e)->id, streamId(s)); return 0; }
cudaError_t cudaEventSynchronize(void* e) { LOG("{\"op\":\"event_sync\",\"event\":%d", static_cast<Event*>(e)->id); return 0; }
cudaError_t cudaEventElapsedTime(float* ms, void* a, void* b)
{
	// what a real device rejects: events created with cudaEventDisableTiming (flag 2), events never recorded
	const Event* ea = static_cast<Event*>(a); const Event* eb = static_cast<Event*>(b);
	if ((ea->flags & 2u) || (eb->flags & 2u) || !ea->recorded || !eb->recorded) { LOG("{\"op\":\"api_error\",\"what\":\"cudaEventElapsedTime on an event without timing or never recorded\""); return 400; }
	*ms = 0.01f;
	return 0;
}
cudaError_t cudaStreamWaitEvent(void* s, void* e, unsigned) { LOG("{\"op\":\"stream_wait\",\"stream\":%d,\"event\":%d", streamId(s), static_cast<Event*>(e)->id); return 0; }

// ---- launches
cudaError_t cudaLaunchKernel(const void* func, dim3 grid, dim3 block, void** args, size_t, void* stream)
{
	logLaunch(func, grid, block, args, stream, 0);
	return 0;
}
// cudaLaunchConfig_t (driver_types.h): gridDim, blockDim, dynamicSmemBytes, stream, attrs, numAttrs
struct LaunchAttr { int id; char pad[4]; union { int programmaticStreamSerializationAllowed; char bytes[64]; } val; };
struct LaunchConfig { dim3 grid, block; size_t shmem; void* stream; LaunchAttr* attrs; unsigned numAttrs; };
cudaError_t cudaLaunchKernelExC(const LaunchConfig* cfg, const void* func, void** args)
{
	int pdl = 0;
	for (unsigned i = 0; i < cfg->numAttrs; ++i)
		if (cfg->attrs[i].id == 6 /* cudaLaunchAttributeProgrammaticStreamSerialization */ && cfg->attrs[i].val.programmaticStreamSerializationAllowed) pdl = 1;
	logLaunch(func, cfg->grid, cfg->block, args, cfg->stream, pdl);
	return 0;
}

// ---- scenario marker: a host-side barrier between the ranks (what torch.distributed.barrier() is in a real group)
void b2r_trace_host_barrier(void) { LOG("{\"op\":\"host_barrier\""); }

// ---- IPC between the contexts of THIS process: the handle carries the address, opening it yields the same memory
struct IpcHandle { char reserved[64]; };
cudaError_t cudaIpcGetMemHandle(IpcHandle* h, void* p) { if (!h || !p) return 1; memset(h, 0, sizeof *h); memcpy(h->reserved, &p, sizeof p); return 0; }
cudaError_t cudaIpcOpenMemHandle(void** p, IpcHandle h, unsigned) { memcpy(p, h.reserved, sizeof *p); return *p ? 0 : 1; }
cudaError_t cudaIpcCloseMemHandle(void*) { return 0; }

}
