// libb2r.so - C ABI (include/b2r.h) and host orchestration of the B200 rendering path.
//
// One frame = one fixed sequence of kernels on one stream, no host synchronisation and no
// device->host round trip inside it (the reference needs ceil(visible/256) batches with a
// cudaDeviceSynchronize after every launch, CUDARenderer.cu:237-330):
//
//   H2D   per-mesh constants (WVP computed on the host exactly like CUDARenderer.cu:242)
//   K1    vertex_transform_kernel      IVertex -> pos4 + attr
//   K2    triangle_setup_kernel        cull / frustum / screen / bbox / bin rect / tile counts
//   K3    bin_big_count_kernel         cooperative counting of large rectangles
//   K4    tile_scan_kernel             exclusive scan of the tile counts (chained over a few CTAs)
//   K5    bin_scatter_kernel           ids -> lists (small rectangles per primitive, large ones cooperatively)
//   K6-8  bin_sort_{runs,mid,long}     canonical (ascending id) order per tile
//   K9    raster_shade_kernel          coverage + depth resolve + deferred shade + ROP write-back
#include "../../include/b2r.h"
#include "device_types.cuh"
#include "geometry_kernels.cuh"
#include "raster_kernel.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace b2r;

namespace {

constexpr int kRing = 4;            // (measured at 8 ranks, round 2: 8 slots / rings of 6 and 8 frames are SLOWER - 7 700 / 5 100 / 1 050 frames/s at depth
constexpr int kMaxFrameRing = 4;    //  4 / 6 / 8: more frames in flight means more resident kernels spinning on peer flags; profiles/r02_ab_log.md)      // B2R_OPT_FRAME_RING: at most this many frames with their own streams / buffer sets (<= kRing)
constexpr int kInfoWords = 16;
constexpr size_t kArenaFlagBytes = 512;   // flag words of the sharded front end: 32 + 8 per ring slot (device_types.cuh: PeerFlags)
#ifndef B2R_DIRECT_MAX_AREA
#define B2R_DIRECT_MAX_AREA 16
#endif
constexpr int kDirectMaxAreaDefault = B2R_DIRECT_MAX_AREA;
thread_local std::string g_createError;

// grow-only device array
template <typename T>
struct DeviceVec
{
	T* ptr = nullptr;
	size_t cap = 0;
	cudaError_t reserve(size_t n, cudaStream_t s, size_t keep = 0)
	{
		if (n <= cap) return cudaSuccess;
		size_t newCap = std::max(n, cap + cap / 2);
		T* np = nullptr;
		cudaError_t e = cudaMalloc((void**)&np, newCap * sizeof(T));
		if (e != cudaSuccess) return e;
		if (keep && ptr)
		{
			e = cudaMemcpyAsync(np, ptr, keep * sizeof(T), cudaMemcpyDeviceToDevice, s);
			if (e == cudaSuccess) e = cudaStreamSynchronize(s);
			if (e != cudaSuccess) { cudaFree(np); return e; }
		}
		cudaFree(ptr);
		ptr = np; cap = newCap;
		return cudaSuccess;
	}
	void release() { cudaFree(ptr); ptr = nullptr; cap = 0; }
};

struct HostMesh
{
	uint32_t vtxInBase, vtxOutBase, nVertices, idxBase, nTris, primBase;
	float world[16];
	int texIds[4];
};

struct HostTexture { uint32_t* px = nullptr; uint32_t w = 0, h = 0; bool alive = false; uint32_t tiledShift = 0; };

struct RenderCall
{
	float camPos[3], camFwd[3], viewProj[16], clearRGB[3];
	int cull, sample, vis;
	bool valid = false;
};

}

struct b2r_context
{
	int device = 0;
	uint32_t width = 0, height = 0;
	uint32_t binsX = 0, binsY = 0, binW = kTile, binH = kTile;
	cudaStream_t ownStream = nullptr, stream = nullptr;
	std::string error;

	// options
	int fillTopLeft = 1, ownerRank = 0, ownerCount = 1, debugSurfaces = 0, stageTimers = 0, sortBins = 0, packedOutput = 0;
	int directMaxArea = kDirectMaxAreaDefault;      // B2R_OPT_DIRECT_MAX_AREA
	bool lastDirect = false;                        // the newest enqueued frame resolved small triangles in its setup kernel (its lists hold the others only)

	// scene
	std::vector<HostMesh> meshes;
	std::vector<HostTexture> textures;
	size_t nVtxIn = 0, nVtxOut = 0, nIdx = 0, nPrims = 0;
	bool sceneDirty = true, texDirty = true;
	DeviceVec<float2> dIVertex;        // 7 float2 per vertex
	DeviceVec<uint32_t> dIndices;
	DeviceVec<float4> dPosIn;          // positions in (x,y,z,1)
	// transformed positions (ndc.xyz, clip.w) and per-mesh constants, double-buffered by frame parity: the vertex
	// stage of frame f+1 runs on its own stream while frame f is still binning / rasterising
	DeviceVec<float4> dPos4[kMaxFrameRing];
	int geoIdx = 0;                    // parity (frame ring: ring slot) of the newest enqueued frame
	std::vector<uint32_t> vsChunkStart; // host copy of the vertex chunks' first transformed-vertex index (ascending)
	cudaStream_t vtxStream = nullptr;  // vertex stage (constants + transform)
	cudaStream_t frontStream = nullptr; // pipelined frames: setup, binning and sort
	cudaEvent_t vtxDone[2] = { nullptr, nullptr }, frontDone[2] = { nullptr, nullptr }, rasterDone[2] = { nullptr, nullptr }, forkEv = nullptr;
	bool rasterPending[2] = { false, false }, vtxNeedsFork = true;
	int overlapVertex = 1, usePdl = 1;
	DeviceVec<float> dOvertexScratch;
	DeviceVec<uint4> dTriRefs;         // static {v0, v1, v2, mesh} per primitive, rebuilt when the mesh list changes
	DeviceVec<float4> dColourOut;      // static vertex colours, indexed like the transformed positions (build_colour_stream_kernel)
	DeviceVec<Chunk> dVsChunks, dTriChunks;
	uint32_t nVsChunks = 0, nTriChunks = 0;
	DeviceVec<MeshConst> dMeshConst[kMaxFrameRing];
	DeviceVec<MeshStatic> dMeshStatic;
	bool meshStaticDirty = true, anyDiffuse = false;
	DeviceVec<TexDesc> dTexDesc;

	// binning
	// Everything the geometry front end (setup -> bin -> sort) writes for one frame.  Two sets: with
	// B2R_OPT_OVERLAP_FRONT the front end of frame f+1 runs on the second stream while frame f is still being
	// rasterised from the other set (set = frame parity, like pos4 / the per-mesh constants); otherwise only set 0
	// is ever allocated.
	struct BinSet
	{
		DeviceVec<uint2> rects;
		DeviceVec<uint32_t> tileCount, tileOffset, tileCursor, longTiles, bigList, binIds;
		DeviceVec<uint4> tileOrder;                   // the raster kernel's work order: one list of {tile, count, offset, key-buffer flag} per length class
		// direct path (B2R_OPT_DIRECT_MAX_AREA): 64-bit depth|id key per pixel in 16x16 blocks, all empty between frames (the
		// raster kernel resets what it merges), and one "holds fragments" byte per tile
		DeviceVec<unsigned long long> keys;
		DeviceVec<uint8_t> dirty;
		DeviceVec<uint32_t> binSlots;                 // direct path: the primitives that still go into the lists, in 32-entry slots (triangle_setup_kernel)
		uint32_t* counters = nullptr;                 // [0] visible, [1] nBig
		unsigned long long* frameInfo = nullptr;      // [0] pairs [1] overflow [2] unsorted tiles [3] long tiles [4] fragments F
	} bins[kMaxFrameRing];
	int overlapFront = 1;              // B2R_OPT_OVERLAP_FRONT
	// Frame ring (B2R_OPT_FRAME_RING = D >= 2, default off): frame f runs from its constants to its raster on ring
	// stream f % D with its own copy of everything a frame writes, so up to D geometry front ends are in flight;
	// only the raster kernels are chained in frame order (enqueueFrameRing)
	int ringDepth = 0;
	cudaStream_t ringStream[kMaxFrameRing] = {};
	cudaEvent_t ringDone[kMaxFrameRing] = {};     // raster + publish of the newest frame of that ring slot
	cudaEvent_t ringRasterDone[kMaxFrameRing] = {}; // its raster kernel alone
	int ringPrev2 = -1;                // ring slot of the frame before the previous one
	int relaxedChain = 1;              // frame ring: raster kernels of consecutive frames may overlap (B2R_RING_RELAXED=0: strictly chained)
	bool ringNeedsFork[kMaxFrameRing] = {};
	int ringPrev = -1;                 // ring slot of the previous frame while it may still be in flight
	uint32_t* dRangeScratch = nullptr;
	int slicedVertex = 0;              // sharded front end: transform only the vertices of this rank's primitive slice;
	                                   // the raster kernel transforms the vertices of its list entries itself (B2R_SHARD_SLICED_VERTEX)
	int overlapFrontSharded = 1;       // ... also with the sharded front end (B2R_OVERLAP_FRONT_SHARDED=0 switches it off)
	int shardLocalColour = 0;          // B2R_OPT_SHARD_LOCAL_COLOUR: sharded front end, every rank keeps its tiles' colours in its own surface
	// distributed Present: the caller's host frame (e.g. a shared-memory mapping), registered and mapped for this device
	void* hostFrame = nullptr; uint32_t* hostFrameDev = nullptr; uint64_t hostFrameBytes = 0;
	cudaEvent_t listsReady[2] = { nullptr, nullptr };   // sharded + pipelined: the "lists ready" barrier of a frame has completed
	bool listsReadyPending[2] = { false, false };
	bool lastPipelined = false;        // the newest enqueued frame ran its front end on the second stream
	struct HostInfo { unsigned long long frameInfo[kInfoWords]; uint32_t counters[2]; };   // device layout: frameInfo, then counters

	// frames in flight: each enqueued frame owns one slot of pinned host memory + events, so the
	// host can run ahead of the device by kRing frames without a synchronisation
	struct Slot
	{
		HostInfo* hInfo = nullptr;                 // pinned
		float* hViewProj = nullptr;                // pinned: the frame's view-projection matrix (read zero-copy by the device)
		MeshStatic* hStatic = nullptr;             // pinned staging of the camera-independent per-mesh constants
		size_t hStaticCap = 0;
		cudaEvent_t done = nullptr;
		cudaEvent_t ev[10] = {};                    // 0..6 stage boundaries; sharded: 7 after pull, 8 after barrier A, 9 after raster
		bool pending = false, timed = false, exchanged = false;
		uint32_t rasterBlocks = 0;                  // worst-case grid of the frame's raster kernel (0: none launched)
		bool longSkipped = false;                   // the frame was enqueued without the long-list sorter (see longLikely)
	} slots[kRing];
	uint64_t seq = 0;
	HostInfo lastInfo{};                           // info of the newest harvested frame
	// Lists longer than kStage are sorted by a kernel of their own (bin_sort_long_kernel), which costs a launch per frame.
	// It is only enqueued while the newest harvested frame of the scene had such lists (or none has been harvested yet);
	// a frame that turns out to have them after all gets its long lists canonicalised when they are read.
	bool longLikely = true, lastLongSkipped = false;
	size_t overflowNeed = 0;
	// raster grid (launchRaster): work items of the last harvested frame and the worst-case grid they were counted under
	uint32_t rasterItemsHint = 0, rasterHintBlocks = 0, lastRasterBlocks = 0;
	bool rasterTailTouched[16] = {};
	uint32_t rasterHintMinSaving = 8192;      // CTAs the hinted grid must save before the tail launch is worth it (launchRaster)
	double stageAccum[B2R_NUM_STAGES] = {};
	float stageLast[B2R_NUM_STAGES] = {};
	uint64_t stageFrames = 0;
	bool haveStageTimes = false;

	// surfaces
	uint32_t* dFramebuffer = nullptr; int32_t* dDepth = nullptr; float* dZ = nullptr; uint32_t* dWinner = nullptr;
	int32_t* dDepthOf[2] = { nullptr, nullptr };    // integer depth surfaces, paired with the colour surfaces (one allocation)
	uint32_t* boundFramebuffer = nullptr;
	// pipelined Present: two colour surfaces; while one is copied to the host on copyStream the next
	// frame renders into the other
	uint32_t* dColour[2] = { nullptr, nullptr };
	int colourIdx = 0;                 // surface the NEXT frame renders into
	int lastColourIdx = 0;             // surface the newest enqueued frame rendered into
	cudaStream_t copyStream = nullptr;
	cudaEvent_t frameDone = nullptr, copyDone[2] = { nullptr, nullptr };
	bool copyPending[2] = { false, false };

	// sharded front end (b2r_shard_*): double-buffered bin lists + flag words in one IPC-exported arena,
	// rank 0's two colour surfaces in another; peers' arenas / rank 0's surfaces mapped over NVLink
	struct Shard
	{
		bool prepared = false, bound = false;
		int n = 1, self = 0;
		int nSets = 2;                         // list sets in the arena: 2 (frame parity) or the depth of the frame ring
		bool rangeValid = false;               // sliced vertex stage: vertex chunks [chunkLo, chunkHi) cover this rank's primitive slice
		uint32_t chunkLo = 0, chunkHi = 0;
		uint8_t* arena = nullptr;
		size_t tilesPad = 0, listCap = 0, setBytes = 0;
		uint8_t* peerArena[kMaxRanks] = {};
		uint32_t* colourBase = nullptr;        // rank 0: own 2-surface allocation; others: the mapping of it
		bool colourIsPeer = false;
		uint32_t frame = 0;                    // frames rendered since the bind (same on every rank)
		uint32_t waitedFrame = 0;              // rank 0: frames whose "colours landed" flags have been waited for
		unsigned long long* hTimeout = nullptr; // pinned: raised by a colour wait that ran into its time-out
		// arena: kArenaFlagBytes of flag words, then per set {tileCount, tileOffset, inbox count[n], inbox offset[n], binIds, region[n]}
		uint32_t* tileCount(uint8_t* base, int set) const { return reinterpret_cast<uint32_t*>(base + kArenaFlagBytes + (size_t)set * setBytes); }
		uint32_t* tileOffset(uint8_t* base, int set) const { return tileCount(base, set) + tilesPad; }
		uint32_t* inboxCount(uint8_t* base, int set, int src) const { return tileCount(base, set) + (size_t)(2 + src) * tilesPad; }
		uint32_t* inboxOffset(uint8_t* base, int set, int src) const { return tileCount(base, set) + (size_t)(2 + n + src) * tilesPad; }
		uint32_t* binIds(uint8_t* base, int set) const { return tileCount(base, set) + (size_t)(2 + 2 * n) * tilesPad; }
		uint32_t* region(uint8_t* base, int set, int src) const { return binIds(base, set) + (size_t)(1 + src) * listCap; }   // lists pushed by rank src
		// direct path (direct = true: decided at b2r_shard_prepare, the same on every rank): behind the list regions of a set, the
		// key-buffer blocks rank src pushed for the tiles owned here (256 keys per owned tile) and its per-tile bytes
		bool direct = false;
		size_t ownedMax = 0;
		unsigned long long* keyInbox(uint8_t* base, int set, int src) const
		{ return reinterpret_cast<unsigned long long*>(region(base, set, n)) + (size_t)src * ownedMax * 256; }
		uint8_t* dirtyInbox(uint8_t* base, int set, int src) const
		{ return reinterpret_cast<uint8_t*>(keyInbox(base, set, n)) + (size_t)src * tilesPad; }
	} shard;
	// bin buffers of the newest frame (the DeviceVecs, or one set of the shard arena)
	uint32_t* curTileCount = nullptr; uint32_t* curTileOffset = nullptr; uint32_t* curBinIds = nullptr;
	const uint2* curRects = nullptr;
	BinSet* curSet = nullptr;          // set of the newest frame
	bool curSorted = true;             // its lists are in canonical order (sorted in-frame, or on the first b2r_read_bins)

	// frame state
	RenderCall last;
	bool haveFrame = false;
	int countFragments = 0;
	uint64_t kernelsLaunched = 0;
	int numSMs = 148;
};

namespace {

int fail(b2r_context* c, int code, const char* fmt, ...)
{
	char buf[512];
	va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
	if (c) c->error = buf; else g_createError = buf;
	return code;
}
#define CU(c, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail((c), B2R_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } while (0)

uint32_t nBins(const b2r_context* c) { return c->binsX * c->binsY; }

// cudaMemset runs on the legacy default stream and may return before the fill has happened; the context's streams are
// non-blocking, i.e. NOT ordered behind the legacy stream.  Allocation-time fills are therefore completed before anything
// else can touch the new buffer (found by tools/trace/stream_order.py).
cudaError_t finishFills(b2r_context*) { return cudaStreamSynchronize(nullptr); }

int allocSurfaces(b2r_context* c)
{
	const size_t n = (size_t)c->width * c->height;
	// two colour surfaces in ONE allocation (pipelined Present; one IPC handle for the sharded mode)
	CU(c, cudaMalloc((void**)&c->dFramebuffer, 2 * n * 4));
	CU(c, cudaMalloc((void**)&c->dDepth, 2 * n * 4));
	c->dDepthOf[0] = c->dDepth; c->dDepthOf[1] = c->dDepth + n;
	CU(c, cudaMemset(c->dFramebuffer, 0, 2 * n * 4));     // CUDAWindowHelper.h:17-25 zero-initialises
	c->dColour[0] = c->dFramebuffer; c->dColour[1] = c->dFramebuffer + n;
	CU(c, cudaMemset(c->dDepth, 0, 2 * n * 4));
	CU(c, finishFills(c));
	return B2R_OK;
}

int ensureDebugSurfaces(b2r_context* c)
{
	const size_t n = (size_t)c->width * c->height;
	const bool fresh = !c->dZ || !c->dWinner;
	if (!c->dZ) { CU(c, cudaMalloc((void**)&c->dZ, n * 4)); CU(c, cudaMemset(c->dZ, 0, n * 4)); }
	if (!c->dWinner) { CU(c, cudaMalloc((void**)&c->dWinner, n * 4)); CU(c, cudaMemset(c->dWinner, 0xFF, n * 4)); }
	if (fresh) CU(c, finishFills(c));
	return B2R_OK;
}

// (Re)size one set of front-end buffers for the current bin grid and scene.  Grow-only and a no-op when nothing
// changed, so enqueueFrame calls it for the set a frame is about to use (the second set exists only once a
// pipelined frame has been enqueued).  Growing frees the old allocation, which synchronises the device.
int reserveBinSet(b2r_context* c, b2r_context::BinSet& B)
{
	const size_t n = nBins(c);
	if (!B.frameInfo)
	{
		CU(c, cudaMalloc((void**)&B.frameInfo, kInfoWords * 8 + kCounterWords * 4));
		CU(c, cudaMemset(B.frameInfo, 0, kInfoWords * 8 + kCounterWords * 4));
		CU(c, finishFills(c));
		B.counters = reinterpret_cast<uint32_t*>(B.frameInfo + kInfoWords);
	}
	// tile_scan_kernel reads the counts with 16-byte loads; tileOffset[n] = overflow marker for peer ranks
	CU(c, B.tileCount.reserve(n + 4, c->stream));
	CU(c, B.tileOffset.reserve(n + 4, c->stream));
	CU(c, B.tileCursor.reserve(n, c->stream));
	CU(c, B.longTiles.reserve(n, c->stream));
	CU(c, B.tileOrder.reserve(n * kTileClasses, c->stream));
	CU(c, B.rects.reserve(std::max<size_t>(c->nPrims, 1), c->stream));
	CU(c, B.bigList.reserve(std::max<size_t>(c->nPrims, 1), c->stream));
	// list capacity heuristic; regrown on overflow (see validateFrame), and kept equal across the sets
	// (a few large triangles can fill every tile: a margin in proportion to the bin grid - 160 entries per tile - keeps e.g.
	// 64 full-screen layers at 4K, 4.15 M pairs, clear of the 4 M floor)
	size_t listCap = std::max<size_t>(std::max<size_t>((size_t)1 << 22, c->nPrims * 3), n * 160);
	for (const auto& other : c->bins) listCap = std::max(listCap, other.binIds.cap);
	CU(c, B.binIds.reserve(listCap, c->stream));
	if (c->directMaxArea)
	{
		const size_t keyWords = (size_t)((c->width + 15) / 16) * ((c->height + 15) / 16) * 256;
		const unsigned long long* oldKeys = B.keys.ptr; const uint8_t* oldDirty = B.dirty.ptr;
		CU(c, B.keys.reserve(keyWords, c->stream));
		CU(c, B.dirty.reserve(n + 16, c->stream));
		CU(c, B.binSlots.reserve(((c->nPrims + 31) / 32 + 8) * 32, c->stream));
		if (B.keys.ptr != oldKeys) CU(c, cudaMemsetAsync(B.keys.ptr, 0xFF, B.keys.cap * 8, c->stream));
		if (B.dirty.ptr != oldDirty) CU(c, cudaMemsetAsync(B.dirty.ptr, 0, B.dirty.cap, c->stream));
		if (B.keys.ptr != oldKeys || B.dirty.ptr != oldDirty) CU(c, cudaStreamSynchronize(c->stream));
	}
	return B2R_OK;
}

int setBins(b2r_context* c, uint32_t bx, uint32_t by, uint32_t bw, uint32_t bh)
{
	if (!bx || !by || !bw || !bh) return fail(c, B2R_E_INVALID, "bin grid and bin size must be non-zero");
	if ((uint64_t)bx * bw < c->width || (uint64_t)by * bh < c->height)
		return fail(c, B2R_E_INVALID, "bins (%ux%u of %ux%u) must cover the %ux%u screen", bx, by, bw, bh, c->width, c->height);
	if (bx > 32767 || by > 32767) return fail(c, B2R_E_INVALID, "too many bins");
	// tile_scan_kernel publishes one slice sum per 2048 tiles into a fixed block of the counters (kScanMaxCtas slices)
	if ((uint64_t)bx * by > (uint64_t)kScanMaxCtas * kScanPerCta)
		return fail(c, B2R_E_INVALID, "%ux%u bins = %llu tiles exceed the %llu the tile scan is sized for", bx, by,
			(unsigned long long)bx * by, (unsigned long long)kScanMaxCtas * kScanPerCta);
	c->binsX = bx; c->binsY = by; c->binW = bw; c->binH = bh;
	c->longLikely = true;
	return reserveBinSet(c, c->bins[0]);
}

bool texValid(const b2r_context* c, int id) { return id >= 0 && (size_t)id < c->textures.size() && c->textures[id].alive; }

// chunk tables + stream sizes after the mesh list changed
int rebuildScene(b2r_context* c)
{
	std::vector<Chunk> vs, tri;
	for (size_t m = 0; m < c->meshes.size(); ++m)
	{
		const HostMesh& hm = c->meshes[m];
		for (uint32_t f = 0; f < hm.nVertices; f += kChunk)
			vs.push_back(Chunk{ (uint32_t)m, std::min<uint32_t>(kChunk, hm.nVertices - f), hm.vtxInBase + f, hm.vtxOutBase + f, 0u, 0u });
		for (uint32_t f = 0; f < hm.nTris; f += kChunk)
			tri.push_back(Chunk{ (uint32_t)m, std::min<uint32_t>(kChunk, hm.nTris - f), hm.idxBase + 3u * f, hm.vtxOutBase, hm.primBase + f, 0u });
	}
	c->nVsChunks = (uint32_t)vs.size(); c->nTriChunks = (uint32_t)tri.size();
	c->vsChunkStart.resize(vs.size());
	for (size_t i = 0; i < vs.size(); ++i) c->vsChunkStart[i] = vs[i].outStart;
	c->shard.rangeValid = false;
	CU(c, c->dVsChunks.reserve(std::max<size_t>(vs.size(), 1), c->stream));
	CU(c, c->dTriChunks.reserve(std::max<size_t>(tri.size(), 1), c->stream));
	if (!vs.empty()) CU(c, cudaMemcpyAsync(c->dVsChunks.ptr, vs.data(), vs.size() * sizeof(Chunk), cudaMemcpyHostToDevice, c->stream));
	if (!tri.empty()) CU(c, cudaMemcpyAsync(c->dTriChunks.ptr, tri.data(), tri.size() * sizeof(Chunk), cudaMemcpyHostToDevice, c->stream));
	CU(c, cudaStreamSynchronize(c->stream));   // vs/tri are host temporaries
	CU(c, c->dTriRefs.reserve(std::max<size_t>(c->nPrims, 1), c->stream));
	if (c->nTriChunks)
	{
		build_trirefs_kernel<<<c->nTriChunks, kChunk, 0, c->stream>>>(c->dTriChunks.ptr, c->dIndices.ptr, c->dTriRefs.ptr);
		CU(c, cudaGetLastError());
	}
	CU(c, c->dColourOut.reserve(std::max<size_t>(c->nVtxOut, 1), c->stream));
	if (c->nVsChunks)
	{
		build_colour_stream_kernel<<<c->nVsChunks, kChunk, 0, c->stream>>>(c->dVsChunks.ptr, c->dIVertex.ptr, c->dColourOut.ptr);
		CU(c, cudaGetLastError());
	}
	for (int i = 0; i < std::max(2, c->ringDepth); ++i)
	{
		CU(c, c->dPos4[i].reserve(std::max<size_t>(c->nVtxOut, 1), c->stream));
		CU(c, c->dMeshConst[i].reserve(std::max<size_t>(c->meshes.size(), 1), c->stream));
	}
	for (auto& B : c->bins)
		if (B.frameInfo) { int r = reserveBinSet(c, B); if (r) return r; }
	c->sceneDirty = false;
	c->longLikely = true;
	return B2R_OK;
}

int rebuildTextures(b2r_context* c)
{
	std::vector<TexDesc> td(std::max<size_t>(c->textures.size(), 1), TexDesc{ nullptr, 0, 0, 0, 0 });
	for (size_t i = 0; i < c->textures.size(); ++i)
		if (c->textures[i].alive) td[i] = TexDesc{ c->textures[i].px, c->textures[i].w, c->textures[i].h, c->textures[i].tiledShift, 0u };
	CU(c, c->dTexDesc.reserve(td.size(), c->stream));
	CU(c, cudaMemcpyAsync(c->dTexDesc.ptr, td.data(), td.size() * sizeof(TexDesc), cudaMemcpyHostToDevice, c->stream));
	CU(c, cudaStreamSynchronize(c->stream));
	c->texDirty = false;
	return B2R_OK;
}

// the direct path of the setup kernel (sharded front end: only when the arena was prepared for it)
bool directActive(const b2r_context* c) { return c->directMaxArea > 0 && (!c->shard.bound || c->shard.direct); }

// the setup kernel instantiation of a frame: fill rule, direct path
cudaError_t launchSetup(const b2r_context* c, cudaStream_t s, b2r_context::BinSet& B, const float4* pos4, uint32_t* tileCount, uint32_t primLo, uint32_t primHi, const FrameParams& fp);

FrameParams makeParams(const b2r_context* c, const RenderCall& rc)
{
	FrameParams fp{};
	for (int i = 0; i < 3; ++i) { fp.camPos[i] = rc.camPos[i]; fp.camFwd[i] = rc.camFwd[i]; fp.clearRGB[i] = rc.clearRGB[i]; }
	fp.width = c->width; fp.height = c->height;
	fp.binsX = c->binsX; fp.binsY = c->binsY; fp.binW = c->binW; fp.binH = c->binH;
	fp.subX = (c->binW + kTile - 1) / kTile; fp.subY = (c->binH + kTile - 1) / kTile;
	fp.binsXMagic = fp.binsX > 1 ? (uint32_t)(0x100000000ull / fp.binsX) : 0u;
	fp.subXMagic = fp.subX > 1 ? (uint32_t)(0x100000000ull / fp.subX) : 0u;
	auto log2OrMinus1 = [](uint32_t v) { int s = 0; while ((1u << s) < v) ++s; return (1u << s) == v ? s : -1; };
	fp.binShiftX = log2OrMinus1(c->binW); fp.binShiftY = log2OrMinus1(c->binH);
	fp.cullMode = (uint32_t)rc.cull;
	// the sign-only back-face test (geometry_kernels.cuh backFace) is proven for camera-forward components of at most 2^20
	if (std::fabs(rc.camFwd[0]) <= 1048576.f && std::fabs(rc.camFwd[1]) <= 1048576.f && std::fabs(rc.camFwd[2]) <= 1048576.f) fp.cullMode |= kCullFastOk;
	fp.sampleState = (uint32_t)rc.sample; fp.visState = (uint32_t)rc.vis;
	fp.fillTopLeft = (uint32_t)c->fillTopLeft;
	fp.ownerRank = (uint32_t)c->ownerRank; fp.ownerCount = (uint32_t)c->ownerCount;
	fp.debugSurfaces = (uint32_t)c->debugSurfaces;
	fp.bigRectTiles = 64;
	fp.nMeshes = (uint32_t)c->meshes.size();
	fp.packedOutput = (uint32_t)c->packedOutput;
	fp.shardGeometry = c->shard.bound ? 1u : 0u;
	const uint32_t subs = fp.subX * fp.subY;
	fp.subsMagic = subs > 1 ? (uint32_t)(0x100000000ull / subs) : 0u;
	// with several 16x16 blocks per bin the CTAs of a bin share its list: it is sorted by a kernel of its own before the raster
	fp.sortLists = (c->sortBins && subs == 1) ? 1u : 0u;
	fp.directMaxArea = directActive(c) ? (uint32_t)c->directMaxArea : 0u;
	fp.keyBlocksX = (c->width + 15) / 16;
	fp.countFragments = (uint32_t)c->countFragments;
	return fp;
}

// kernel launch with the programmatic-dependent-launch attribute (see pdlPrologue in device_types.cuh)
template <typename... P, typename... A>
cudaError_t launchFrameKernel(const b2r_context* c, void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, A&&... args)
{
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
	cudaLaunchAttribute attr{};
	attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr.val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = &attr; cfg.numAttrs = c->usePdl ? 1u : 0u;
	return cudaLaunchKernelEx(&cfg, kernel, std::forward<A>(args)...);
}

// the raster kernel instantiation of a frame: fill rule, fragment counter, light variant (the frame cannot reach the
// texture / visualiser paths)
// `blocks` = the 16x16 pixel blocks of the frame, an upper bound of the kernel's work items (CTAs behind the last item leave at once)
// The raster kernel + its tail (raster_kernel.cuh).  `blocks` = the worst case, every tile of this rank a work item; the
// main grid follows the item count of the last finished frame of the same tile grid (rasterItemsHint, harvestSlot), the
// tail's CTAs stride over the items behind it.  Without a hint - first frames, after a change of the tile grid - the
// main grid is the worst case and there is no tail.
cudaError_t launchRaster(b2r_context* c, uint32_t blocks, cudaStream_t s, const RasterArgs& A, const FrameParams& fp, bool textured, bool otf = false)
{
	const int variant = (c->fillTopLeft ? 1 : 0) | (c->countFragments ? 2 : 0) | (textured ? 4 : 0) | (otf ? 8 : 0);
	const dim3 block(kTile * kTile);
	uint32_t mainBlocks = blocks;
	// B2R_RASTER_GRID_HINT=0 in the environment: always the worst-case grid (A/B and debugging aid)
	static const bool useHint = [] { const char* e = getenv("B2R_RASTER_GRID_HINT"); return !e || atoi(e) != 0; }();
	if (useHint && c->rasterHintBlocks == blocks && c->rasterItemsHint)
	{
		// Only where it was measured to pay: the tail is one more launch in the frame's chain (~5 us by itself), so the
		// smaller grid has to save more than that.  4K config 3 (8 927 CTAs fewer): raster 116 -> 109 us.  1080p vehicle
		// (6 500 fewer): 50.3 -> 50.7 us, a wash.  640x480 vehicle (640 fewer, main grid below one wave): the stage took
		// 80 us instead of 37 - not understood, and no GPU time left to look into it - hence the floor of 8 192 CTAs
		// (b2r_context::rasterHintMinSaving; B2R_RASTER_HINT_MIN_SAVING in the environment of b2r_create overrides it: the tests
		// exercise the tail on small frames).
		const uint32_t hinted = std::min<uint32_t>(blocks, c->rasterItemsHint + c->rasterItemsHint / 4u + 256u);
		if (blocks - hinted >= c->rasterHintMinSaving) mainBlocks = hinted;
	}
	const dim3 grid(mainBlocks);
	c->lastRasterBlocks = blocks;
	cudaError_t e = cudaErrorInvalidValue;
	if (!c->rasterTailTouched[variant])
	{
		// First raster launch of this variant: query its tail kernel, so that CUDA's lazy loading brings the function in now
		// and not with the first hinted frame - which comes a few frames later, possibly inside somebody's timed region
		// (a load stalls the host for a millisecond or two).
		int n = 0;
#define B2R_TOUCH_CASE(v, tl, cnt, tex) case v: (void)cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, raster_tail_kernel<tl, cnt, tex, false>, (int)(kTile * kTile), 0); break; \
	case v + 8: (void)cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, raster_tail_kernel<tl, cnt, tex, true>, (int)(kTile * kTile), 0); break;
		switch (variant)
		{
		B2R_TOUCH_CASE(0, false, false, false) B2R_TOUCH_CASE(1, true, false, false)
		B2R_TOUCH_CASE(2, false, true, false)  B2R_TOUCH_CASE(3, true, true, false)
		B2R_TOUCH_CASE(4, false, false, true)  B2R_TOUCH_CASE(5, true, false, true)
		B2R_TOUCH_CASE(6, false, true, true)   B2R_TOUCH_CASE(7, true, true, true)
		default: break;
		}
#undef B2R_TOUCH_CASE
		(void)cudaGetLastError();
		c->rasterTailTouched[variant] = true;
	}
#define B2R_RASTER_CASE(v, tl, cnt, tex) case v: e = launchFrameKernel(c, raster_shade_kernel<tl, cnt, tex, false>, grid, block, 0, s, A, fp); break; \
	case v + 8: e = launchFrameKernel(c, raster_shade_kernel<tl, cnt, tex, true>, grid, block, 0, s, A, fp); break;
	switch (variant)
	{
	B2R_RASTER_CASE(0, false, false, false) B2R_RASTER_CASE(1, true, false, false)
	B2R_RASTER_CASE(2, false, true, false)  B2R_RASTER_CASE(3, true, true, false)
	B2R_RASTER_CASE(4, false, false, true)  B2R_RASTER_CASE(5, true, false, true)
	B2R_RASTER_CASE(6, false, true, true)   B2R_RASTER_CASE(7, true, true, true)
	default: break;
	}
#undef B2R_RASTER_CASE
	++c->kernelsLaunched;
	if (e != cudaSuccess || mainBlocks == blocks) return e;
	const dim3 tailGrid(std::min<uint32_t>(blocks - mainBlocks, (uint32_t)c->numSMs * 2u));
#define B2R_TAIL_CASE(v, tl, cnt, tex) case v: e = launchFrameKernel(c, raster_tail_kernel<tl, cnt, tex, false>, tailGrid, block, 0, s, A, fp, mainBlocks); break; \
	case v + 8: e = launchFrameKernel(c, raster_tail_kernel<tl, cnt, tex, true>, tailGrid, block, 0, s, A, fp, mainBlocks); break;
	switch (variant)
	{
	B2R_TAIL_CASE(0, false, false, false) B2R_TAIL_CASE(1, true, false, false)
	B2R_TAIL_CASE(2, false, true, false)  B2R_TAIL_CASE(3, true, true, false)
	B2R_TAIL_CASE(4, false, false, true)  B2R_TAIL_CASE(5, true, false, true)
	B2R_TAIL_CASE(6, false, true, true)   B2R_TAIL_CASE(7, true, true, true)
	default: break;
	}
#undef B2R_TAIL_CASE
	++c->kernelsLaunched;
	return e;
}

// persistent grid of the setup kernel: as many CTAs as are resident at once (each strides over the 256-primitive chunks)
template <typename K>
uint32_t setupGrid(const b2r_context* c, K kernel, uint32_t prims, int& cached)
{
	if (cached < 1)
	{
		int n = 0;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kChunk, 0) != cudaSuccess || n < 1) { (void)cudaGetLastError(); n = B2R_SETUP_CTAS; }
		cached = n;
	}
	uint32_t grid = std::max<uint32_t>(1u, std::min<uint32_t>((uint32_t)c->numSMs * (uint32_t)cached, (prims + kChunk - 1) / kChunk));
	// testing aid: a smaller grid makes every CTA walk several chunks (the copy pipeline's steady state) on small scenes too
	static const int forced = [] { const char* e = getenv("B2R_SETUP_GRID"); return e ? atoi(e) : 0; }();
	if (forced > 0) grid = std::min<uint32_t>(grid, (uint32_t)forced);
	return grid;
}

cudaError_t launchSetup(const b2r_context* c, cudaStream_t s, b2r_context::BinSet& B, const float4* pos4, uint32_t* tileCount, uint32_t primLo, uint32_t primHi, const FrameParams& fp)
{
	static thread_local int occ[4] = { 0, 0, 0, 0 };      // resident CTAs per SM of the four instantiations (a property of the binary)
	const dim3 block(kChunk);
#define B2R_SETUP_CASE(i, tl, direct) return launchFrameKernel(c, triangle_setup_kernel<tl, direct>, dim3(setupGrid(c, triangle_setup_kernel<tl, direct>, primHi - primLo, occ[i])), block, 0, s, \
		(const uint4*)c->dTriRefs.ptr, pos4, B.rects.ptr, tileCount, B.bigList.ptr, B.counters, B.binSlots.ptr, B.keys.ptr, B.dirty.ptr, B.frameInfo, primLo, primHi, fp)
	if (fp.directMaxArea) { if (c->fillTopLeft) B2R_SETUP_CASE(0, true, true); B2R_SETUP_CASE(1, false, true); }
	if (c->fillTopLeft) B2R_SETUP_CASE(2, true, false);
	B2R_SETUP_CASE(3, false, false);
#undef B2R_SETUP_CASE
}

// sharded front end + direct path: this rank's touched key-buffer blocks go to the tiles' owners (shard_push_keys_kernel)
int enqueueKeyPush(b2r_context* c, b2r_context::BinSet& B, int set, cudaStream_t s)
{
	const b2r_context::Shard& sh = c->shard;
	ShardKeyPush P{};
	for (int p = 0; p < sh.n; ++p)
	{
		if (p == sh.self) continue;
		P.keyInbox[p] = sh.keyInbox(sh.peerArena[p], set, sh.self);
		P.dirtyInbox[p] = sh.dirtyInbox(sh.peerArena[p], set, sh.self);
	}
	const uint32_t tiles = nBins(c);
	CU(c, launchFrameKernel(c, shard_push_keys_kernel, dim3(std::min<uint32_t>((tiles + 7) / 8, (uint32_t)c->numSMs * 2u)), dim3(256), 0, s, B.keys.ptr, B.dirty.ptr, P, tiles, (uint32_t)sh.self, (uint32_t)sh.n));
	++c->kernelsLaunched;
	return B2R_OK;
}

#define STAGE_EVENT(c, slot, i) do { if ((slot).timed) CU((c), cudaEventRecord((slot).ev[(i)], (c)->stream)); } while (0)

// book-keeping of a finished frame: capacity overflow, stage times, counters
int harvestSlot(b2r_context* c, b2r_context::Slot& slot)
{
	slot.pending = false;
	c->lastInfo = *slot.hInfo;
	c->longLikely = slot.hInfo->frameInfo[3] != 0;
	c->lastLongSkipped = slot.longSkipped;
	if (slot.hInfo->frameInfo[1])
		c->overflowNeed = std::max(c->overflowNeed, (size_t)slot.hInfo->frameInfo[0]);
	else if (slot.rasterBlocks)
	{
		// work items of that frame's raster kernel: the blocks of the non-empty classes + the batches of empty ones
		const uint32_t subs = ((c->binW + kTile - 1) / kTile) * ((c->binH + kTile - 1) / kTile);
		unsigned long long items = 0;
		for (uint32_t k = 0; k + 1 < kTileClasses; ++k) items += slot.hInfo->frameInfo[kInfoClassBase + k] * subs;
		items += (slot.hInfo->frameInfo[kInfoClassBase + kTileClasses - 1] * subs + kEmptyBatch - 1) / kEmptyBatch;
		c->rasterItemsHint = (uint32_t)std::min<unsigned long long>(items, slot.rasterBlocks);
		c->rasterHintBlocks = slot.rasterBlocks;
	}
	if (slot.hInfo->frameInfo[7])
		return fail(c, B2R_E_STATE, "sharded front end: a peer rank did not reach the frame barrier within 4 s");
	if (slot.timed)
	{
		float total = 0.f;
		for (int i = 0; i < 6; ++i)
		{
			float ms = 0.f;
			CU(c, cudaEventElapsedTime(&ms, slot.ev[i], slot.ev[i + 1]));
			c->stageLast[i] = ms; c->stageAccum[i] += ms; total += ms;
		}
		float ms = 0.f;
		CU(c, cudaEventElapsedTime(&ms, slot.ev[0], slot.ev[6]));
		c->stageLast[B2R_STAGE_TOTAL] = ms; c->stageAccum[B2R_STAGE_TOTAL] += ms;
		float pull = 0.f, wait = 0.f;
		if (slot.exchanged)
		{
			float waitA = 0.f, raster = 0.f, waitB = 0.f;
			CU(c, cudaEventElapsedTime(&waitA, slot.ev[5], slot.ev[8]));
			CU(c, cudaEventElapsedTime(&pull, slot.ev[8], slot.ev[7]));
			CU(c, cudaEventElapsedTime(&raster, slot.ev[7], slot.ev[9]));
			CU(c, cudaEventElapsedTime(&waitB, slot.ev[9], slot.ev[6]));
			wait = waitA + waitB;
			c->stageAccum[B2R_STAGE_RASTER_SHADE] += raster - c->stageLast[B2R_STAGE_RASTER_SHADE];
			c->stageLast[B2R_STAGE_RASTER_SHADE] = raster;
		}
		c->stageLast[B2R_STAGE_EXCHANGE] = pull; c->stageAccum[B2R_STAGE_EXCHANGE] += pull;
		c->stageLast[B2R_STAGE_PEER_WAIT] = wait; c->stageAccum[B2R_STAGE_PEER_WAIT] += wait;
		++c->stageFrames;
		c->haveStageTimes = true;
	}
	return B2R_OK;
}

// camera-independent per-mesh constants (MeshStatic): uploaded only when the mesh list, a world matrix or the texture
// set changed, through the frame slot's pinned staging buffer, on the caller's stream `s`.  The streams that run the
// constants kernels must be ordered behind this copy: vtxNeedsFork.
int uploadMeshStatic(b2r_context* c, b2r_context::Slot& slot, cudaStream_t s)
{
	if (slot.hStaticCap < c->meshes.size())
	{
		if (slot.hStatic) cudaFreeHost(slot.hStatic);
		slot.hStatic = nullptr; slot.hStaticCap = 0;
		const size_t cap = std::max<size_t>(c->meshes.size() * 2, 16);
		CU(c, cudaMallocHost((void**)&slot.hStatic, cap * sizeof(MeshStatic)));
		slot.hStaticCap = cap;
	}
	c->anyDiffuse = false;
	for (size_t m = 0; m < c->meshes.size(); ++m)
	{
		const HostMesh& hm = c->meshes[m];
		MeshStatic& ms = slot.hStatic[m];
		memcpy(ms.world, hm.world, 64);
		ms.vtxInBase = hm.vtxInBase; ms.vtxOutBase = hm.vtxOutBase; ms.nVertices = hm.nVertices;
		ms.idxBase = hm.idxBase; ms.nTris = hm.nTris; ms.primBase = hm.primBase;
		// CUDARenderer.cu:559-571: Norm only with Diff, Spec/Gloss only with Norm
		const bool d = texValid(c, hm.texIds[0]), n = d && texValid(c, hm.texIds[1]);
		ms.tex[0] = d ? hm.texIds[0] : -1;
		ms.tex[1] = n ? hm.texIds[1] : -1;
		ms.tex[2] = (n && texValid(c, hm.texIds[2])) ? hm.texIds[2] : -1;
		ms.tex[3] = (n && texValid(c, hm.texIds[3])) ? hm.texIds[3] : -1;
		ms.pad[0] = ms.pad[1] = 0;
		c->anyDiffuse |= d;
	}
	CU(c, c->dMeshStatic.reserve(std::max<size_t>(c->meshes.size(), 1), s));
	if (!c->meshes.empty())
		CU(c, cudaMemcpyAsync(c->dMeshStatic.ptr, slot.hStatic, c->meshes.size() * sizeof(MeshStatic), cudaMemcpyHostToDevice, s));
	c->vtxNeedsFork = true;
	c->meshStaticDirty = false;
	return B2R_OK;
}

// Canonical (ascending primitive id) order of the tile lists.  Inside a frame the lists of up to kStage entries are
// sorted by the raster kernel itself (the CTA that rasterises a tile stages its list in shared memory anyway); what is
// left for a kernel of its own are the few longer lists (tile_scan_kernel's "long" work list), before the raster.
// everything = true (b2r_read_bins after frames rendered with B2R_OPT_SORT_BINS = 0, or with the sharded front end,
// where a tile's list is sorted at its owner): also the short lists, one CTA per tile.
int enqueueListSort(b2r_context* c, b2r_context::BinSet& B, uint32_t* tileCount, uint32_t* tileOffset, uint32_t* binIds, cudaStream_t s, bool everything)
{
	const uint32_t tiles = nBins(c);
	if (everything)
	{
		bin_sort_tiles_kernel<<<std::min<uint32_t>(tiles, (uint32_t)c->numSMs * 6u), kSortThreads, 0, s>>>(tileCount, tileOffset, binIds, tiles, B.frameInfo);
		CU(c, cudaGetLastError());
	}
	CU(c, launchFrameKernel(c, bin_sort_long_kernel, dim3(c->numSMs), dim3(1024), kSortLongSmem, s, B.longTiles.ptr, tileCount, tileOffset,
		binIds, B.frameInfo));
	return B2R_OK;
}

int enqueueFrameRing(b2r_context* c);

int enqueueFrame(b2r_context* c)
{
	if (c->ringDepth) return enqueueFrameRing(c);
	const RenderCall& rc = c->last;
	if (c->sceneDirty) { int r = rebuildScene(c); if (r) return r; }
	if (c->texDirty) { int r = rebuildTextures(c); if (r) return r; }
	if (c->debugSurfaces) { int r = ensureDebugSurfaces(c); if (r) return r; }
	cudaStream_t s = c->stream;
	const FrameParams fp = makeParams(c, rc);
	const uint32_t tiles = nBins(c);
	c->kernelsLaunched = 0;

	// Pipelined front end (B2R_OPT_OVERLAP_FRONT): the whole geometry front end of this frame (constants, vertex,
	// setup, binning, sort) goes to the second stream and works on the bin set of the frame's parity, so it
	// overlaps the raster kernel of the previous frame, which reads the other set.  Not while the stages are being
	// timed and not with the sharded front end (its sets live in the peer-visible arena).
	const bool pipelined = c->overlapFront && c->overlapVertex && !c->stageTimers && c->nVsChunks && (!c->shard.bound || c->overlapFrontSharded);
	if (pipelined != c->lastPipelined)
	{
		// frames of the other kind used set 0 / the main stream regardless of parity: order this frame's front end
		// behind everything enqueued so far
		c->vtxNeedsFork = true;
		c->lastPipelined = pipelined;
	}
	b2r_context::BinSet& B = c->bins[pipelined ? (int)(c->seq & 1u) : 0];
	{ int r = reserveBinSet(c, B); if (r) return r; }

	// bin buffers and primitive slice of this frame
	const b2r_context::Shard& sh = c->shard;
	const int set = (int)(sh.frame & 1u);
	uint32_t* tileCount = sh.bound ? sh.tileCount(sh.arena, set) : B.tileCount.ptr;
	uint32_t* tileOffset = sh.bound ? sh.tileOffset(sh.arena, set) : B.tileOffset.ptr;
	uint32_t* binIds = sh.bound ? sh.binIds(sh.arena, set) : B.binIds.ptr;
	const size_t listCap = sh.bound ? sh.listCap : B.binIds.cap;
	c->curTileCount = tileCount; c->curTileOffset = tileOffset; c->curBinIds = binIds; c->curRects = B.rects.ptr;
	// sharded front end: a tile's list is put in order at its owner; the rank's own buffer is canonicalised when it is read
	c->curSet = &B; c->curSorted = c->sortBins != 0 && !sh.bound;
	uint32_t primLo = 0, primHi = (uint32_t)c->nPrims;
	if (sh.bound)
	{
		// contiguous slices of whole 256-primitive blocks, so that the concatenation of the ranks' lists in
		// rank order is the ascending list a single GPU builds
		const uint64_t blocks = (c->nPrims + kChunk - 1) / kChunk;
		primLo = (uint32_t)std::min<uint64_t>(c->nPrims, blocks * (uint64_t)sh.self / (uint64_t)sh.n * kChunk);
		primHi = (uint32_t)std::min<uint64_t>(c->nPrims, blocks * (uint64_t)(sh.self + 1) / (uint64_t)sh.n * kChunk);
	}

	b2r_context::Slot& slot = c->slots[c->seq % kRing];
	if (slot.pending)
	{
		CU(c, cudaEventSynchronize(slot.done));
		int r = harvestSlot(c, slot); if (r) return r;
	}
	slot.timed = c->stageTimers != 0;
	c->lastRasterBlocks = 0;
	slot.exchanged = slot.timed && c->shard.bound;
	memcpy(slot.hViewProj, rc.viewProj, 64);
	if (c->meshStaticDirty) { int r = uploadMeshStatic(c, slot, s); if (r) return r; }
	STAGE_EVENT(c, slot, 0);
	const uint32_t nMeshes = (uint32_t)c->meshes.size();
	const int g = (int)(c->seq & 1u);
	c->geoIdx = g;
	float4* pos4 = c->dPos4[g].ptr;
	MeshConst* meshConst = c->dMeshConst[g].ptr;
	// The vertex stage depends on nothing of the previous frame, so (unless the stages are being timed) it runs on
	// its own stream: vertex(f+1) overlaps binning / raster of frame f.  It re-uses the buffers of frame f-1, hence
	// the wait for that frame's raster; after an idle point it is ordered behind whatever the caller enqueued.
	const bool overlap = c->overlapVertex && !slot.timed && c->nVsChunks;
	if (overlap)
	{
		cudaStream_t v = c->vtxStream;
		if (c->vtxNeedsFork)
		{
			CU(c, cudaEventRecord(c->forkEv, s));
			CU(c, cudaStreamWaitEvent(v, c->forkEv, 0));
			CU(c, cudaStreamWaitEvent(c->frontStream, c->forkEv, 0));
			c->vtxNeedsFork = false;
		}
		if (c->rasterPending[g]) CU(c, cudaStreamWaitEvent(v, c->rasterDone[g], 0));
		CU(c, launchFrameKernel(c, build_constants_kernel, dim3(std::max<uint32_t>(1u, std::min<uint32_t>((nMeshes + 63) / 64, (uint32_t)c->numSMs * 4))), dim3(64), 0, v, 
			slot.hViewProj, c->dMeshStatic.ptr, meshConst, nMeshes));
		CU(c, launchFrameKernel(c, vertex_transform_kernel, dim3((c->nVsChunks + kVsPerCta - 1) / kVsPerCta), dim3(kChunk), 0, v, c->dVsChunks.ptr, c->nVsChunks, meshConst, c->dPosIn.ptr, pos4));
		c->kernelsLaunched += 2;
		CU(c, cudaEventRecord(c->vtxDone[g], v));
	}
	// Stream of the front end (setup, binning, sort): the main stream, or - pipelined - a stream of its own, joined
	// into the main stream before the raster.  Three frames are then in flight: vertex(f+2) | front end(f+1) |
	// raster(f), each stage waiting only for what it reads or recycles.
	cudaStream_t fs = pipelined ? c->frontStream : s;
	if (overlap) CU(c, cudaStreamWaitEvent(fs, c->vtxDone[g], 0));
	if (pipelined)
	{
		// this front end recycles the bin set of frame f-2 (read by that frame's raster and publish kernels) ...
		if (c->rasterPending[g]) CU(c, cudaStreamWaitEvent(fs, c->rasterDone[g], 0));
		// ... and, sharded, stores list heads / lists into the peers' set of this parity, which their raster kernels
		// of frame f-2 read.  A peer has finished that raster before it arrives at the "lists ready" barrier of
		// frame f-1 (stream order), so once OUR barrier of frame f-1 has completed every peer's set is free.
		if (c->shard.bound && c->listsReadyPending[g ^ 1]) CU(c, cudaStreamWaitEvent(fs, c->listsReady[g ^ 1], 0));
	}
	// one small kernel instead of three memsets: memsets may be served by a copy engine, where they would
	// queue behind the previous frame's 33 MB Present copy
	const uint32_t clearMeshes = overlap ? 0u : nMeshes;
	const uint32_t clearCtas = std::min<uint32_t>(std::max<uint32_t>((tiles + 4095) / 4096, (clearMeshes + 255) / 256), (uint32_t)c->numSMs * 4);
	CU(c, launchFrameKernel(c, frame_clear_kernel, dim3(clearCtas), dim3(256), 0, fs, tileCount, tiles, B.counters, B.frameInfo, (uint32_t)kInfoWords,
		kScanBase + 2u * ((tiles + kScanPerCta - 1) / kScanPerCta),
		slot.hViewProj, c->dMeshStatic.ptr, meshConst, clearMeshes, PeerFlags{}, 0u, 0u, 0u));
	++c->kernelsLaunched;

	if (c->nVsChunks && !overlap)
	{
		CU(c, launchFrameKernel(c, vertex_transform_kernel, dim3((c->nVsChunks + kVsPerCta - 1) / kVsPerCta), dim3(kChunk), 0, fs, c->dVsChunks.ptr, c->nVsChunks, meshConst, c->dPosIn.ptr, pos4));
		++c->kernelsLaunched;
	}
	STAGE_EVENT(c, slot, 1);
	if (c->nTriChunks)
	{
		if (primHi > primLo)
		{
			CU(c, launchSetup(c, fs, B, pos4, tileCount, primLo, primHi, fp));
			++c->kernelsLaunched;
			if (sh.bound && fp.directMaxArea) { int r = enqueueKeyPush(c, B, set, fs); if (r) return r; }
		}
		CU(c, launchFrameKernel(c, bin_big_count_kernel, dim3(c->numSMs * 4), dim3(256), 0, fs, B.bigList.ptr, B.counters, B.rects.ptr, tileCount, fp));
		++c->kernelsLaunched;
	}
	STAGE_EVENT(c, slot, 2);
	{
		// sharded: the heads of this rank's lists go to the tiles' owners from the scan kernel itself, so the transfer hides
		// behind scatter + push of the lists
		ShardInbox inbox{};
		if (sh.bound)
			for (int p = 0; p < sh.n; ++p)
			{
				uint8_t* base = (p == sh.self) ? sh.arena : sh.peerArena[p];
				inbox.tileCount[p] = sh.inboxCount(base, set, sh.self); inbox.tileOffset[p] = sh.inboxOffset(base, set, sh.self);
			}
		CU(c, launchFrameKernel(c, tile_scan_kernel, dim3((tiles + kScanPerCta - 1) / kScanPerCta), dim3(kScanThreads), 0, fs, tileCount, tileOffset, B.tileCursor.ptr, tiles,
			(unsigned long long)listCap, (uint32_t)kStage, B.longTiles.ptr, sh.bound ? nullptr : B.tileOrder.ptr, fp.directMaxArea ? (const uint8_t*)B.dirty.ptr : (const uint8_t*)nullptr,
			(uint32_t)c->ownerRank, (uint32_t)c->ownerCount, B.counters, inbox, sh.bound ? (uint32_t)sh.n : 0u, B.frameInfo));
		++c->kernelsLaunched;
	}
	STAGE_EVENT(c, slot, 3);
	if (c->nPrims)
	{
		{
			// per-primitive CTAs, followed by the CTAs that scatter the cooperatively binned (big) triangles
			// (direct path: a fixed number of CTAs strides over the slots of the primitives that are binned at all)
			const uint32_t allBlocks = (primHi - primLo + 256 * kScatterPerThread - 1) / (256 * kScatterPerThread);
			const uint32_t primBlocks = fp.directMaxArea ? std::max<uint32_t>(1u, std::min<uint32_t>(allBlocks, (uint32_t)c->numSMs * 4u)) : allBlocks;
			// sharded: the ids of tiles other ranks own are mirrored into those ranks' arenas as they are scattered
			ShardRegions R{};
			if (sh.bound)
				for (int p = 0; p < sh.n; ++p) R.region[p] = (p == sh.self) ? nullptr : sh.region(sh.peerArena[p], set, sh.self);
			CU(c, launchFrameKernel(c, bin_scatter_kernel, dim3(primBlocks + (uint32_t)c->numSMs * 2), dim3(256), 0, fs, B.rects.ptr, fp.directMaxArea ? (const uint32_t*)B.binSlots.ptr : (const uint32_t*)nullptr, primLo, primHi, tileOffset,
				B.tileCursor.ptr, binIds, B.bigList.ptr, B.counters, primBlocks, B.frameInfo, fp, R, sh.bound ? (uint32_t)sh.n : 0u, (uint32_t)sh.self));
			++c->kernelsLaunched;
		}
	}
	STAGE_EVENT(c, slot, 4);
	const bool bigBins = c->binW > (uint32_t)kTile || c->binH > (uint32_t)kTile;
	slot.longSkipped = c->sortBins && c->nPrims && !bigBins && !c->longLikely;
	if (c->sortBins && c->nPrims && !slot.longSkipped)
	{
		int r = enqueueListSort(c, B, tileCount, tileOffset, binIds, fs, bigBins); if (r) return r;
		c->kernelsLaunched += bigBins ? 2 : 1;
	}
	STAGE_EVENT(c, slot, 5);
	if (pipelined)
	{
		// join: the raster of this frame (main stream) starts when its lists are sorted
		CU(c, cudaEventRecord(c->frontDone[g], fs));
		CU(c, cudaStreamWaitEvent(s, c->frontDone[g], 0));
	}
	{
		RasterArgs A{};
		A.tileCount = tileCount; A.tileOffset = tileOffset; A.binIds = binIds;
		A.triRefs = c->dTriRefs.ptr; A.pos4 = pos4; A.posIn = c->dPosIn.ptr; A.ivertex = c->dIVertex.ptr; A.colour = c->dColourOut.ptr;
		A.meshes = meshConst; A.textures = c->dTexDesc.ptr; A.frameInfo = B.frameInfo; A.tileOrder = B.tileOrder.ptr;
		A.keys = B.keys.ptr; A.dirty = B.dirty.ptr; c->lastDirect = fp.directMaxArea != 0u;
		if (sh.bound) c->colourIdx = set;       // every rank writes rank 0's surface `frame & 1`
		if (!c->boundFramebuffer && c->copyPending[c->colourIdx])
		{
			// this surface is still being copied to the host by an earlier b2r_present_async
			CU(c, cudaStreamWaitEvent(s, c->copyDone[c->colourIdx], 0));
			c->copyPending[c->colourIdx] = false;
		}
		A.framebuffer = c->boundFramebuffer ? c->boundFramebuffer : c->dColour[c->colourIdx];
		c->lastColourIdx = c->colourIdx;
		PeerFlags pf{};
		const uint32_t owned = (tiles > (uint32_t)c->ownerRank) ? (tiles - (uint32_t)c->ownerRank + (uint32_t)c->ownerCount - 1) / (uint32_t)c->ownerCount : 0;
		if (sh.bound)
		{
			ShardTable T{};
			for (int p = 0; p < sh.n; ++p)
			{
				uint8_t* base = (p == sh.self) ? sh.arena : sh.peerArena[p];
				T.tileCount[p] = sh.inboxCount(sh.arena, set, p); T.tileOffset[p] = sh.inboxOffset(sh.arena, set, p);
				T.binIds[p] = (p == sh.self) ? sh.binIds(sh.arena, set) : sh.region(sh.arena, set, p);
				if (sh.direct) { T.keyInbox[p] = sh.keyInbox(sh.arena, set, p); T.dirtyInbox[p] = sh.dirtyInbox(sh.arena, set, p); }
				pf.flags[p] = reinterpret_cast<uint32_t*>(base);
			}
			if (!c->shardLocalColour) A.framebuffer = sh.colourBase + (size_t)set * c->width * c->height;
			// "lists ready": nobody reads a peer's lists before that peer has sorted them.  (Rank 0 arrives only
			// after the Present copy of this surface - the wait above - so no peer overwrites pixels in flight.)
			CU(c, launchFrameKernel(c, peer_barrier_kernel, dim3(1), dim3(32), 0, s, pf, (uint32_t)sh.self, (uint32_t)sh.n, 0u, sh.frame + 1u, 1u, 1u, B.frameInfo + 7, 1u));
			++c->kernelsLaunched;
			CU(c, cudaEventRecord(c->listsReady[g], s));
			c->listsReadyPending[g] = true;
			STAGE_EVENT(c, slot, 8);
			A.shard = T;
			// the raster kernel's work order from the merged heads: this rank's tiles, longest lists first
			if (owned)
			{
				CU(c, launchFrameKernel(c, shard_merge_heads_kernel, dim3((owned + 255) / 256), dim3(256), 0, s, T, tiles, (uint32_t)sh.self, (uint32_t)sh.n, B.tileOrder.ptr, B.frameInfo,
					fp.directMaxArea ? (const uint8_t*)B.dirty.ptr : (const uint8_t*)nullptr));
				++c->kernelsLaunched;
			}
			STAGE_EVENT(c, slot, 7);
		}
		A.depth = c->dDepthOf[c->colourIdx]; A.zSurface = c->dZ; A.winnerSurface = c->dWinner;
		if (owned)
		{
			// light variant when the frame cannot reach the texture / visualiser paths
			CU(c, launchRaster(c, owned * fp.subX * fp.subY, s, A, fp, rc.vis != 0 || c->anyDiffuse));
		}
		// the vertex stage of frame f+2 may overwrite pos4 / constants of this parity once this raster is done
		// (pipelined: the whole front end of frame f+2 re-uses this bin set, whose counters publish_info_kernel
		// still reads - the event is recorded after it, below)
		if (!pipelined)
		{
			CU(c, cudaEventRecord(c->rasterDone[g], s));
			c->rasterPending[g] = true;
		}
		if (sh.bound)
		{
			STAGE_EVENT(c, slot, 9);
			++c->shard.frame;
		}
	}
	// "colours landed": publish_info_kernel tells rank 0 that this rank's tiles of the frame are in rank 0's surface;
	// rank 0 waits for the flags only when it reads or presents the frame (enqueueColourWait)
	uint32_t* landedFlag = nullptr;
	if (sh.bound && !c->shardLocalColour)
		landedFlag = reinterpret_cast<uint32_t*>(sh.self == 0 ? sh.arena : sh.peerArena[0]) + 8 + sh.self;
	STAGE_EVENT(c, slot, 6);
	// frameInfo and counters[0..1] are adjacent on the device (one allocation): one copy
	CU(c, launchFrameKernel(c, publish_info_kernel, dim3(1), dim3(32), 0, s, B.frameInfo, reinterpret_cast<unsigned long long*>(slot.hInfo), (uint32_t)(sizeof(b2r_context::HostInfo) / 8),
		landedFlag, c->shard.frame, PeerFlags{}, 0u, 0u));
	++c->kernelsLaunched;
	if (pipelined)
	{
		CU(c, cudaEventRecord(c->rasterDone[g], s));
		c->rasterPending[g] = true;
	}
	CU(c, cudaEventRecord(slot.done, s));
	CU(c, cudaGetLastError());
	slot.pending = true;
	slot.rasterBlocks = c->lastRasterBlocks;
	++c->seq;
	c->haveFrame = true;
	return B2R_OK;
}

// Frame ring (B2R_OPT_FRAME_RING = D): the alternative to the three-stage pipeline of enqueueFrame.  Frame f runs
// from its constants to its publish kernel on ring stream f % D and owns ring slot f % D of everything a frame
// writes (transformed positions, per-mesh constants, bin set and - sharded - the list set in the arenas), so
//   * re-use of a slot is ordered by the stream itself (frame f follows frame f-D on the same stream);
//   * the geometry front ends of up to D consecutive frames are in flight at once, which is what a rank of the
//     sharded front end needs: its front end is a chain of short, latency-bound kernels (1/n of the primitives)
//     that cannot fill the GPU, and in the three-stage pipeline the front end of frame f+1 has to wait for the
//     cross-GPU barrier of frame f;
//   * only the raster kernels are chained in frame order (event of the previous frame): they share the depth and
//     colour surfaces, and the "colours landed" / "frame done" epochs must rise monotonically.
// Cross-GPU re-use of a list set: rank A's front end of frame f stores list heads and lists into set f % D of every
// peer, which the peer's raster kernel of frame f-D reads.  Every rank therefore announces "frame done" to all
// ranks from its publish kernel (flag words 16..23), and A waits for all peers' announcements of frame f-D right
// before its first store into peer memory (normally long satisfied: D-1 frames lie in between).
// With B2R_OPT_STAGE_TIMERS a frame waits for its predecessor at its START, so the stage times are those of
// serial frames (all stages of a frame are on one stream here).
int enqueueFrameRing(b2r_context* c)
{
	const RenderCall& rc = c->last;
	if (c->sceneDirty) { int r = rebuildScene(c); if (r) return r; }
	if (c->texDirty) { int r = rebuildTextures(c); if (r) return r; }
	if (c->debugSurfaces) { int r = ensureDebugSurfaces(c); if (r) return r; }
	cudaStream_t s = c->stream;
	const FrameParams fp = makeParams(c, rc);
	const uint32_t tiles = nBins(c);
	c->kernelsLaunched = 0;
	b2r_context::Shard& sh = c->shard;
	const int D = c->ringDepth;
	if (sh.bound && sh.nSets != D) return fail(c, B2R_E_STATE, "frame ring of depth %d but the shared arena holds %d list sets", D, sh.nSets);
	// ring slot: the same on every rank of a sharded group (all ranks render the same frames after the bind)
	const uint32_t frameNo = sh.bound ? sh.frame : (uint32_t)c->seq;
	const int k = (int)(frameNo % (uint32_t)D);
	cudaStream_t rs = c->ringStream[k];
	b2r_context::BinSet& B = c->bins[k];
	{ int r = reserveBinSet(c, B); if (r) return r; }
	c->lastPipelined = false;

	uint32_t* tileCount = sh.bound ? sh.tileCount(sh.arena, k) : B.tileCount.ptr;
	uint32_t* tileOffset = sh.bound ? sh.tileOffset(sh.arena, k) : B.tileOffset.ptr;
	uint32_t* binIds = sh.bound ? sh.binIds(sh.arena, k) : B.binIds.ptr;
	const size_t listCap = sh.bound ? sh.listCap : B.binIds.cap;
	c->curTileCount = tileCount; c->curTileOffset = tileOffset; c->curBinIds = binIds; c->curRects = B.rects.ptr;
	// sharded front end: a tile's list is put in order at its owner; the rank's own buffer is canonicalised when it is read
	c->curSet = &B; c->curSorted = c->sortBins != 0 && !sh.bound;
	uint32_t primLo = 0, primHi = (uint32_t)c->nPrims;
	if (sh.bound)
	{
		const uint64_t blocks = (c->nPrims + kChunk - 1) / kChunk;
		primLo = (uint32_t)std::min<uint64_t>(c->nPrims, blocks * (uint64_t)sh.self / (uint64_t)sh.n * kChunk);
		primHi = (uint32_t)std::min<uint64_t>(c->nPrims, blocks * (uint64_t)(sh.self + 1) / (uint64_t)sh.n * kChunk);
	}
	// sliced vertex stage (sharded only): this rank transforms only the vertex chunks its primitive slice refers to; its
	// raster kernel (OTF instantiation) transforms the vertices of its list entries and of its pixels' winners itself
	const bool sliced = sh.bound && c->slicedVertex;
	if (sliced && !sh.rangeValid)
	{
		sh.chunkLo = sh.chunkHi = 0;
		if (primHi > primLo && c->nVsChunks)
		{
			uint32_t range[2] = { 0xFFFFFFFFu, 0u };
			if (!c->dRangeScratch) CU(c, cudaMalloc((void**)&c->dRangeScratch, 8));
			uint32_t* dRange = c->dRangeScratch;
			CU(c, cudaMemcpyAsync(dRange, range, 8, cudaMemcpyHostToDevice, s));
			prim_vertex_range_kernel<<<std::min<uint32_t>((primHi - primLo + 255) / 256, (uint32_t)c->numSMs * 8), 256, 0, s>>>(c->dTriRefs.ptr, primLo, primHi, dRange);
			CU(c, cudaGetLastError());
			CU(c, cudaMemcpyAsync(range, dRange, 8, cudaMemcpyDeviceToHost, s));
			CU(c, cudaStreamSynchronize(s));
			if (range[0] <= range[1])
			{
				// chunks are in ascending order of their first transformed-vertex index
				const auto& cs = c->vsChunkStart;
				sh.chunkLo = (uint32_t)(std::upper_bound(cs.begin(), cs.end(), range[0]) - cs.begin()) - 1u;
				sh.chunkHi = (uint32_t)(std::upper_bound(cs.begin(), cs.end(), range[1]) - cs.begin());
			}
		}
		sh.rangeValid = true;
	}

	b2r_context::Slot& slot = c->slots[c->seq % kRing];
	if (slot.pending)
	{
		CU(c, cudaEventSynchronize(slot.done));
		int r = harvestSlot(c, slot); if (r) return r;
	}
	slot.timed = c->stageTimers != 0;
	c->lastRasterBlocks = 0;
	slot.exchanged = slot.timed && sh.bound;
	memcpy(slot.hViewProj, rc.viewProj, 64);
	// (the copy goes to the caller's stream, which has joined every frame enqueued so far - their constants kernels have
	// read the old table; the ring streams fork from it below)
	if (c->meshStaticDirty) { int r = uploadMeshStatic(c, slot, s); if (r) return r; }
	if (c->vtxNeedsFork)
	{
		// after an idle point (or the copy above): every ring stream is ordered behind what the caller's stream holds now
		CU(c, cudaEventRecord(c->forkEv, s));
		for (int i = 0; i < D; ++i) c->ringNeedsFork[i] = true;
		c->vtxNeedsFork = false;
	}
	if (c->ringNeedsFork[k])
	{
		CU(c, cudaStreamWaitEvent(rs, c->forkEv, 0));
		c->ringNeedsFork[k] = false;
	}
	bool waitedPrev = false;
	if (slot.timed && c->ringPrev >= 0 && c->ringPrev != k)
	{
		CU(c, cudaStreamWaitEvent(rs, c->ringDone[c->ringPrev], 0));      // serial frames while the stages are timed
		waitedPrev = true;
	}
#define RING_STAGE_EVENT(i) do { if (slot.timed) CU(c, cudaEventRecord(slot.ev[(i)], rs)); } while (0)
	RING_STAGE_EVENT(0);
	PeerFlags pf{};
	if (sh.bound)
		for (int p = 0; p < sh.n; ++p) pf.flags[p] = reinterpret_cast<uint32_t*>(p == sh.self ? sh.arena : sh.peerArena[p]);
	const uint32_t nMeshes = (uint32_t)c->meshes.size();
	c->geoIdx = k;
	float4* pos4 = c->dPos4[k].ptr;
	MeshConst* meshConst = c->dMeshConst[k].ptr;
	{
		// clear + constants, then the vertex stage (all chunks, or the chunks of this rank's primitive slice)
		const uint32_t clearCtas = std::min<uint32_t>(std::max<uint32_t>((tiles + 4095) / 4096, (nMeshes + 255) / 256), (uint32_t)c->numSMs * 4);
		CU(c, launchFrameKernel(c, frame_clear_kernel, dim3(std::max<uint32_t>(clearCtas, 1u)), dim3(256), 0, rs, tileCount, tiles, B.counters, B.frameInfo, (uint32_t)kInfoWords,
			kScanBase + 2u * ((tiles + kScanPerCta - 1) / kScanPerCta),
			slot.hViewProj, c->dMeshStatic.ptr, meshConst, nMeshes,
			pf, (uint32_t)sh.self, (sh.bound && sh.frame >= (uint32_t)D) ? (uint32_t)sh.n : 0u, sh.frame - (uint32_t)D + 1u));
		++c->kernelsLaunched;
		const uint32_t cLo = sliced ? sh.chunkLo : 0u, cHi = sliced ? sh.chunkHi : c->nVsChunks;
		if (cHi > cLo)
		{
			CU(c, launchFrameKernel(c, vertex_transform_kernel, dim3((cHi - cLo + kVsPerCta - 1) / kVsPerCta), dim3(kChunk), 0, rs, c->dVsChunks.ptr + cLo, cHi - cLo, meshConst, c->dPosIn.ptr, pos4));
			++c->kernelsLaunched;
		}
	}
	RING_STAGE_EVENT(1);
	if (c->nTriChunks)
	{
		if (primHi > primLo)
		{
			CU(c, launchSetup(c, rs, B, pos4, tileCount, primLo, primHi, fp));
			++c->kernelsLaunched;
			if (sh.bound && fp.directMaxArea) { int r = enqueueKeyPush(c, B, k, rs); if (r) return r; }
		}
		CU(c, launchFrameKernel(c, bin_big_count_kernel, dim3(c->numSMs * 4), dim3(256), 0, rs, B.bigList.ptr, B.counters, B.rects.ptr, tileCount, fp));
		++c->kernelsLaunched;
	}
	RING_STAGE_EVENT(2);
	{
		ShardInbox inbox{};
		if (sh.bound)
			for (int p = 0; p < sh.n; ++p)
			{
				uint8_t* base = (p == sh.self) ? sh.arena : sh.peerArena[p];
				inbox.tileCount[p] = sh.inboxCount(base, k, sh.self); inbox.tileOffset[p] = sh.inboxOffset(base, k, sh.self);
			}
		CU(c, launchFrameKernel(c, tile_scan_kernel, dim3((tiles + kScanPerCta - 1) / kScanPerCta), dim3(kScanThreads), 0, rs, tileCount, tileOffset, B.tileCursor.ptr, tiles,
			(unsigned long long)listCap, (uint32_t)kStage, B.longTiles.ptr, sh.bound ? nullptr : B.tileOrder.ptr, fp.directMaxArea ? (const uint8_t*)B.dirty.ptr : (const uint8_t*)nullptr,
			(uint32_t)c->ownerRank, (uint32_t)c->ownerCount, B.counters, inbox, sh.bound ? (uint32_t)sh.n : 0u, B.frameInfo));
		++c->kernelsLaunched;
	}
	RING_STAGE_EVENT(3);
	if (c->nPrims)
	{
		const uint32_t allBlocks = (primHi - primLo + 256 * kScatterPerThread - 1) / (256 * kScatterPerThread);
		const uint32_t primBlocks = fp.directMaxArea ? std::max<uint32_t>(1u, std::min<uint32_t>(allBlocks, (uint32_t)c->numSMs * 4u)) : allBlocks;
		ShardRegions R{};
		if (sh.bound)
			for (int p = 0; p < sh.n; ++p) R.region[p] = (p == sh.self) ? nullptr : sh.region(sh.peerArena[p], k, sh.self);
		CU(c, launchFrameKernel(c, bin_scatter_kernel, dim3(primBlocks + (uint32_t)c->numSMs * 2), dim3(256), 0, rs, B.rects.ptr, fp.directMaxArea ? (const uint32_t*)B.binSlots.ptr : (const uint32_t*)nullptr, primLo, primHi, tileOffset,
			B.tileCursor.ptr, binIds, B.bigList.ptr, B.counters, primBlocks, B.frameInfo, fp, R, sh.bound ? (uint32_t)sh.n : 0u, (uint32_t)sh.self));
		++c->kernelsLaunched;
	}
	RING_STAGE_EVENT(4);
	const bool bigBins = c->binW > (uint32_t)kTile || c->binH > (uint32_t)kTile;
	slot.longSkipped = c->sortBins && c->nPrims && !bigBins && !c->longLikely;
	if (c->sortBins && c->nPrims && !slot.longSkipped)
	{
		int r = enqueueListSort(c, B, tileCount, tileOffset, binIds, rs, bigBins); if (r) return r;
		c->kernelsLaunched += bigBins ? 2 : 1;
	}
	RING_STAGE_EVENT(5);
	{
		RasterArgs A{};
		A.tileCount = tileCount; A.tileOffset = tileOffset; A.binIds = binIds;
		A.triRefs = c->dTriRefs.ptr; A.pos4 = pos4; A.posIn = c->dPosIn.ptr; A.ivertex = c->dIVertex.ptr; A.colour = c->dColourOut.ptr;
		A.meshes = meshConst; A.textures = c->dTexDesc.ptr; A.frameInfo = B.frameInfo; A.tileOrder = B.tileOrder.ptr;
		A.keys = B.keys.ptr; A.dirty = B.dirty.ptr; c->lastDirect = fp.directMaxArea != 0u;
		// the frames of the ring alternate strictly between the two colour + depth surface pairs (sharded: every rank writes
		// rank 0's surface `frame & 1`), which is what lets consecutive raster kernels overlap (below)
		c->colourIdx = sh.bound ? (int)(sh.frame & 1u) : (int)(c->seq & 1u);
		if (!c->boundFramebuffer && c->copyPending[c->colourIdx])
		{
			// this surface is still being copied to the host by an earlier b2r_present_async (sharded: rank 0 arrives at
			// the barrier below only after that copy, so no peer overwrites pixels in flight)
			CU(c, cudaStreamWaitEvent(rs, c->copyDone[c->colourIdx], 0));
			c->copyPending[c->colourIdx] = false;
		}
		A.framebuffer = c->boundFramebuffer ? c->boundFramebuffer : c->dColour[c->colourIdx];
		c->lastColourIdx = c->colourIdx;
		const uint32_t owned = (tiles > (uint32_t)c->ownerRank) ? (tiles - (uint32_t)c->ownerRank + (uint32_t)c->ownerCount - 1) / (uint32_t)c->ownerCount : 0;
		if (sh.bound)
		{
			ShardTable T{};
			for (int p = 0; p < sh.n; ++p)
			{
				T.tileCount[p] = sh.inboxCount(sh.arena, k, p); T.tileOffset[p] = sh.inboxOffset(sh.arena, k, p);
				T.binIds[p] = (p == sh.self) ? sh.binIds(sh.arena, k) : sh.region(sh.arena, k, p);
				if (sh.direct) { T.keyInbox[p] = sh.keyInbox(sh.arena, k, p); T.dirtyInbox[p] = sh.dirtyInbox(sh.arena, k, p); }
			}
			if (!c->shardLocalColour) A.framebuffer = sh.colourBase + (size_t)(sh.frame & 1u) * c->width * c->height;
			// "lists ready": nobody reads the lists a peer pushed before that peer has sorted and pushed them
			// (one set of flag words per ring slot, 32 + 8 k: the barriers of consecutive frames sit on different ring streams
			// and may run in either order - with shared words the older frame's epoch could overwrite the newer one's, and a
			// newer epoch would release the older frame's barrier before that frame's lists have been pushed; frames of the
			// same slot follow each other on one stream, so a slot's words rise monotonically)
			CU(c, launchFrameKernel(c, peer_barrier_kernel, dim3(1), dim3(32), 0, rs, pf, (uint32_t)sh.self, (uint32_t)sh.n, 32u + 8u * (uint32_t)k, sh.frame + 1u, 1u, 1u, B.frameInfo + 7, 0u));
			++c->kernelsLaunched;
			RING_STAGE_EVENT(8);
			A.shard = T;
			if (owned)
			{
				CU(c, launchFrameKernel(c, shard_merge_heads_kernel, dim3((owned + 255) / 256), dim3(256), 0, rs, T, tiles, (uint32_t)sh.self, (uint32_t)sh.n, B.tileOrder.ptr, B.frameInfo,
					fp.directMaxArea ? (const uint8_t*)B.dirty.ptr : (const uint8_t*)nullptr));
				++c->kernelsLaunched;
			}
			RING_STAGE_EVENT(7);
		}
		A.depth = c->dDepthOf[c->colourIdx]; A.zSurface = c->dZ; A.winnerSurface = c->dWinner;
		// raster kernels in frame order: one depth surface, two colour surfaces, monotonic epochs
		// Frames alternate between two colour + depth surface pairs, so the raster kernel of frame f only has to follow
		// the raster of frame f-2 (same pair); consecutive frames' raster kernels may overlap - a rank's raster kernel is a
		// few waves of latency-bound CTAs that do not fill the GPU on their own.  The publish kernels stay chained in frame
		// order (below): the "colours landed" / "frame done" epochs must rise monotonically.  The debug surfaces (one pair)
		// and the strict mode keep the raster kernels themselves in frame order.
		const bool relaxed = c->relaxedChain && !c->debugSurfaces && !c->boundFramebuffer && !c->packedOutput;
		if (!waitedPrev)
		{
			if (!relaxed) { if (c->ringPrev >= 0 && c->ringPrev != k) CU(c, cudaStreamWaitEvent(rs, c->ringDone[c->ringPrev], 0)); }
			else if (c->ringPrev2 >= 0 && c->ringPrev2 != k) CU(c, cudaStreamWaitEvent(rs, c->ringRasterDone[c->ringPrev2], 0));
		}
		if (owned)
		{
			// sliced vertex stage: this rank has transformed the vertices of its own primitive slice only - the raster kernel
			// transforms the vertices of its list entries and of its pixels' winners itself
			CU(c, launchRaster(c, owned * fp.subX * fp.subY, rs, A, fp, rc.vis != 0 || c->anyDiffuse, sliced));
		}
		if (sh.bound)
		{
			RING_STAGE_EVENT(9);
			++sh.frame;
		}
	}
	uint32_t* landedFlag = nullptr;
	if (sh.bound && !c->shardLocalColour)
		landedFlag = reinterpret_cast<uint32_t*>(sh.self == 0 ? sh.arena : sh.peerArena[0]) + 8 + sh.self;
	CU(c, cudaEventRecord(c->ringRasterDone[k], rs));
	if (c->ringPrev >= 0 && c->ringPrev != k) CU(c, cudaStreamWaitEvent(rs, c->ringDone[c->ringPrev], 0));      // publish kernels in frame order
	RING_STAGE_EVENT(6);
#undef RING_STAGE_EVENT
	CU(c, launchFrameKernel(c, publish_info_kernel, dim3(1), dim3(32), 0, rs, B.frameInfo, reinterpret_cast<unsigned long long*>(slot.hInfo), (uint32_t)(sizeof(b2r_context::HostInfo) / 8),
		landedFlag, sh.frame, pf, sh.bound ? (uint32_t)sh.n : 0u, (uint32_t)sh.self));
	++c->kernelsLaunched;
	CU(c, cudaEventRecord(c->ringDone[k], rs));
	c->ringPrev2 = c->ringPrev;
	c->ringPrev = k;
	CU(c, cudaEventRecord(slot.done, rs));
	// the caller's stream observes the frame (events it records, Present, read-backs) without holding back later frames
	CU(c, cudaStreamWaitEvent(s, c->ringDone[k], 0));
	CU(c, cudaGetLastError());
	slot.pending = true;
	slot.rasterBlocks = c->lastRasterBlocks;
	++c->seq;
	c->haveFrame = true;
	return B2R_OK;
}

// sharded front end, rank 0: before the newest frame is read or presented, wait until every rank's colours of
// it have landed in this rank's surface (flag words 8..15 of the arena, stored by the ranks' publish_info_kernel)
int enqueueColourWait(b2r_context* c)
{
	b2r_context::Shard& sh = c->shard;
	if (!sh.bound || sh.self != 0 || sh.waitedFrame == sh.frame || c->shardLocalColour) return B2R_OK;
	PeerFlags pf{};
	for (int p = 0; p < sh.n; ++p) pf.flags[p] = reinterpret_cast<uint32_t*>(p == sh.self ? sh.arena : sh.peerArena[p]);
	CU(c, launchFrameKernel(c, peer_barrier_kernel, dim3(1), dim3(32), 0, c->stream, pf, (uint32_t)sh.self, (uint32_t)sh.n, 8u, sh.frame, 0u, 1u, sh.hTimeout, 1u));
	sh.waitedFrame = sh.frame;
	return B2R_OK;
}

// wait for every frame in flight; if the bin-list buffer was too small for any of them, grow it
// and run the newest frame again (older frames were overwritten anyway)
int validateFrame(b2r_context* c)
{
	for (int attempt = 0; attempt < 4; ++attempt)
	{
		{ int r = enqueueColourWait(c); if (r) return r; }
		CU(c, cudaStreamSynchronize(c->stream));      // the vertex stream is joined into it every frame
		if (c->shard.hTimeout && *c->shard.hTimeout)
		{
			*c->shard.hTimeout = 0;
			return fail(c, B2R_E_STATE, "sharded front end: a peer rank's colours did not arrive within 4 s");
		}
		c->vtxNeedsFork = true; c->rasterPending[0] = c->rasterPending[1] = false; c->listsReadyPending[0] = c->listsReadyPending[1] = false;
		c->ringPrev = -1; c->ringPrev2 = -1;
		for (uint64_t k = 0; k < kRing; ++k)
		{
			b2r_context::Slot& slot = c->slots[(c->seq + k) % kRing];     // oldest first
			if (slot.pending) { int r = harvestSlot(c, slot); if (r) return r; }
		}
		if (!c->overflowNeed) return B2R_OK;
		if (c->shard.bound)
		{
			const size_t need = c->overflowNeed;
			c->overflowNeed = 0;
			return fail(c, B2R_E_STATE, "sharded front end: this rank's bin lists need %zu entries but the shared arena holds %zu; "
				"b2r_shard_release and b2r_shard_prepare with a larger list_capacity on every rank", need, c->shard.listCap);
		}
		const size_t need = c->overflowNeed;
		c->overflowNeed = 0;
		for (auto& B : c->bins)
			if (B.frameInfo)
			{
				CU(c, B.binIds.reserve(need + need / 4, c->stream));
				// a voided frame's raster kernel did not merge (and reset) what its setup kernel put into the key buffer
				if (B.keys.ptr) CU(c, cudaMemsetAsync(B.keys.ptr, 0xFF, B.keys.cap * 8, c->stream));
				if (B.dirty.ptr) CU(c, cudaMemsetAsync(B.dirty.ptr, 0, B.dirty.cap, c->stream));
			}
		int r = enqueueFrame(c);
		if (r) return r;
	}
	return fail(c, B2R_E_STATE, "bin list capacity could not be satisfied");
}

}

static void releaseShard(b2r_context* c)
{
	b2r_context::Shard& sh = c->shard;
	for (int p = 0; p < kMaxRanks; ++p)
		if (sh.peerArena[p]) { cudaIpcCloseMemHandle(sh.peerArena[p]); sh.peerArena[p] = nullptr; }
	if (sh.colourIsPeer && sh.colourBase) cudaIpcCloseMemHandle(sh.colourBase);
	if (sh.arena) cudaFree(sh.arena);
	if (sh.hTimeout) cudaFreeHost(sh.hTimeout);
	sh = b2r_context::Shard{};
}

// ---- self test: the hoisted division of raster_kernel.cuh against __fdiv_rn ------------------
__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
	x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
	return x;
}
__global__ void division_selftest_kernel(unsigned long long nPairs, uint32_t seed, unsigned long long* mismatches)
{
	unsigned long long bad = 0;
	for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nPairs; i += (unsigned long long)gridDim.x * blockDim.x)
	{
		const uint32_t h0 = hash32((uint32_t)i ^ seed), h1 = hash32(h0 + (uint32_t)(i >> 32) + 0x9E3779B9u), h2 = hash32(h1 ^ 0x85EBCA6Bu);
		// divisor: any normal float with exponent in [-95, 95]
		const uint32_t eb = 127u - 95u + (h0 % 191u);
		const float b = __uint_as_float((eb << 23) | (h1 & 0x7FFFFFu));
		// dividend: the domain the raster kernel uses the hoisted path on: a in [b*2^-64, 4b], a >= 2^-95
		float a;
		if ((h2 & 3u) == 0u) a = b * (1.0001f * (float)((h2 >> 8) + 1u) * (1.f / 16777216.f));            // u uniform in (0, 1.0001]
		else if ((h2 & 3u) == 1u) a = b * __uint_as_float(((127u - (h2 >> 8) % 64u) << 23) | (h0 & 0x7FFFFFu));   // u = 2^-k * m, k < 64
		else if ((h2 & 3u) == 2u) a = b * __uint_as_float(((127u + (h2 >> 8) % 2u) << 23) | (h2 & 0x7FFFFFu));    // u in [1, 4)
		else a = __uint_as_float(__float_as_uint(b) - (h2 >> 8) % 4096u);                                  // a within a few thousand ulp below b
		if (!(a >= fmaxf(kDivLo, b * 5.42101086e-20f)) || !(b >= kDivLo && b <= kDivHi)) continue;
		const float q = divByArea(a, b, refinedReciprocal(b));
		if (__float_as_uint(q) != __float_as_uint(__fdiv_rn(a, b))) ++bad;
	}
	if (bad) atomicAdd(mismatches, bad);
}

// ---- self test: the sign-only back-face test of geometry_kernels.cuh against the reference's arithmetic --------------
// Operands are drawn to sit where the short cut could go wrong: dot products within a few ulp of zero (the normal is
// built perpendicular to the camera axis and then nudged), normals from 2^-70 to 2^70 (both sides of the magnitude
// guard), zero / subnormal components, axis-aligned and non-unit camera axes up to the 2^20 the host admits, and
// NaN / Inf.  out[0] counts disagreements, out[1] the cases the short cut decided.
__global__ void cull_selftest_kernel(unsigned long long nCases, uint32_t seed, unsigned long long* out)
{
	unsigned long long bad = 0, fast = 0;
	for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nCases; i += (unsigned long long)gridDim.x * blockDim.x)
	{
		uint32_t h[8];
		h[0] = hash32((uint32_t)i ^ seed);
		for (int k = 1; k < 8; ++k) h[k] = hash32(h[k - 1] + (uint32_t)(i >> 32) + 0x9E3779B9u * (uint32_t)k + 0x85EBCA6Bu * seed);      // (h[0] alone visits the same 2^32 values for every seed)
		auto unit = [](uint32_t x) { return (float)(x >> 8) * (1.f / 8388608.f) - 1.f; };      // [-1, 1)
		// camera axis
		float c[3] = { unit(h[0]), unit(h[1]), unit(h[2]) };
		const uint32_t cKind = h[3] & 7u;
		if (cKind == 0u) { c[0] = 0.f; c[1] = 0.f; c[2] = (h[3] & 8u) ? 1.f : -1.f; }
		else if (cKind == 1u) c[(h[3] >> 4) % 3u] = 0.f;
		else if (cKind == 2u) { const float k = __uint_as_float((127u + (h[3] >> 4) % 20u) << 23); c[0] *= k; c[1] *= k; c[2] *= k; }
		else if (cKind == 3u) { const float k = __uint_as_float((127u - (h[3] >> 4) % 40u) << 23); c[0] *= k; c[1] *= k; c[2] *= k; }
		else { float x = c[0], y = c[1], z = c[2]; normalize3(x, y, z); c[0] = x; c[1] = y; c[2] = z; }
		// normal
		float n[3] = { unit(h[4]), unit(h[5]), unit(h[6]) };
		const uint32_t nKind = h[7] & 7u;
		if (nKind <= 3u)
		{
			// perpendicular to c up to rounding, then nudged by 0 .. 2^-12 of c
			const float cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2], nc = n[0] * c[0] + n[1] * c[1] + n[2] * c[2];
			const float k = cc > 0.f ? nc / cc : 0.f;
			const float nudge = (nKind == 0u) ? 0.f : unit(h[7] >> 3 | h[6] << 29) * __uint_as_float((127u - 12u - (h[7] >> 8) % 24u) << 23);
			for (int a = 0; a < 3; ++a) n[a] = fmaf(-k + nudge, c[a], n[a]);
		}
		else if (nKind == 4u) n[(h[7] >> 4) % 3u] = 0.f;
		else if (nKind == 5u) { n[(h[7] >> 4) % 3u] = 0.f; n[(h[7] >> 6) % 3u] = 0.f; }
		// magnitude 2^-70 .. 2^70, mostly the 2^-30 .. 2^0 of real cross products
		const int e = (h[7] & 0x100u) ? -(int)((h[7] >> 9) % 31u) : (int)((h[7] >> 9) % 141u) - 70;
		const float scale = __uint_as_float((uint32_t)(127 + e) << 23);
		for (int a = 0; a < 3; ++a) n[a] *= scale;
		if ((h[7] >> 20) % 4096u == 0u) n[(h[7] >> 4) % 3u] = __uint_as_float(0x7FC00000u);         // NaN
		if ((h[7] >> 20) % 4096u == 1u) n[(h[7] >> 4) % 3u] = __uint_as_float(0x7F800000u);         // Inf
		if ((h[7] >> 20) % 4096u == 2u) n[(h[7] >> 4) % 3u] = __uint_as_float(h[6] & 0x7FFFFFu);    // subnormal
		const bool admitted = fabsf(c[0]) <= 1048576.f && fabsf(c[1]) <= 1048576.f && fabsf(c[2]) <= 1048576.f;      // makeParams
		#pragma unroll
		for (uint32_t mode = 0; mode < 2; ++mode)
		{
			const bool a = backFace(n[0], n[1], n[2], c, mode | (admitted ? kCullFastOk : 0u));
			const bool b = backFaceExact(n[0], n[1], n[2], c, mode);
			if (a != b) ++bad;
		}
		bool back;
		if (admitted && backFaceCertain(n[0], n[1], n[2], c, back)) ++fast;
	}
	if (bad) atomicAdd(&out[0], bad);
	if (fast) atomicAdd(&out[1], fast);
}

// =======================================================================================
extern "C" {

// test hook (not part of include/b2r.h): the sign-only back-face test against the exact one on nCases pseudo-random
// adversarial (normal, camera axis) pairs; counts[0] = disagreements, counts[1] = cases the short cut decided
int b2r_selftest_cull(int device, unsigned long long nCases, uint32_t seed, unsigned long long counts[2])
{
	if (!counts) return B2R_E_INVALID;
	if (cudaSetDevice(device) != cudaSuccess) return B2R_E_NO_DEVICE;
	unsigned long long* d = nullptr;
	if (cudaMalloc((void**)&d, 16) != cudaSuccess) return B2R_E_CUDA;
	cudaMemset(d, 0, 16);
	cull_selftest_kernel<<<148 * 8, 256>>>(nCases, seed, d);
	cudaError_t e = cudaMemcpy(counts, d, 16, cudaMemcpyDeviceToHost);
	cudaFree(d);
	return e == cudaSuccess ? B2R_OK : B2R_E_CUDA;
}

// test hook (not part of include/b2r.h): bit-compares the raster kernel's hoisted division with
// __fdiv_rn on nPairs pseudo-random operand pairs; returns the number of mismatches
int b2r_selftest_division(int device, unsigned long long nPairs, uint32_t seed, unsigned long long* mismatches)
{
	if (!mismatches) return B2R_E_INVALID;
	if (cudaSetDevice(device) != cudaSuccess) return B2R_E_NO_DEVICE;
	unsigned long long* d = nullptr;
	if (cudaMalloc((void**)&d, 8) != cudaSuccess) return B2R_E_CUDA;
	cudaMemset(d, 0, 8);
	division_selftest_kernel<<<148 * 8, 256>>>(nPairs, seed, d);
	cudaError_t e = cudaMemcpy(mismatches, d, 8, cudaMemcpyDeviceToHost);
	cudaFree(d);
	return e == cudaSuccess ? B2R_OK : B2R_E_CUDA;
}

const char* b2r_last_error(const b2r_context* ctx)
{
	return ctx ? ctx->error.c_str() : g_createError.c_str();
}

int b2r_create(int device, uint32_t width, uint32_t height, b2r_context** out)
{
	if (!out) return fail(nullptr, B2R_E_INVALID, "out is null");
	*out = nullptr;
	if (!width || !height || width > 32767 || height > 32767)
		return fail(nullptr, B2R_E_INVALID, "resolution %ux%u out of range (bounding boxes are 16-bit, BoundingBox.h:4-10)", width, height);
	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
		return fail(nullptr, B2R_E_NO_DEVICE, "no CUDA device: this renderer has no CPU fallback");
	if (device < 0 || device >= count) return fail(nullptr, B2R_E_INVALID, "device %d out of range (%d devices)", device, count);
	b2r_context* c = new (std::nothrow) b2r_context;
	if (!c) return fail(nullptr, B2R_E_INVALID, "out of host memory");
	c->device = device; c->width = width; c->height = height;
	if (const char* e = getenv("B2R_RASTER_HINT_MIN_SAVING")) c->rasterHintMinSaving = (uint32_t)std::max(0, atoi(e));      // testing aid (launchRaster)
	if (const char* e = getenv("B2R_OVERLAP_VERTEX")) c->overlapVertex = atoi(e) != 0;
	if (const char* e = getenv("B2R_PDL")) c->usePdl = atoi(e) != 0;                           // debugging aid: programmatic dependent launch
	if (const char* e = getenv("B2R_OVERLAP_FRONT")) c->overlapFront = atoi(e) != 0;           // see B2R_OPT_OVERLAP_FRONT
	if (const char* e = getenv("B2R_OVERLAP_FRONT_SHARDED")) c->overlapFrontSharded = atoi(e) != 0;
	if (const char* e = getenv("B2R_FRAME_RING")) { const int d = atoi(e); c->ringDepth = (d >= 2 && d <= kMaxFrameRing) ? d : 0; }   // see B2R_OPT_FRAME_RING
	if (const char* e = getenv("B2R_SHARD_SLICED_VERTEX")) c->slicedVertex = atoi(e) != 0;       // see B2R_OPT_SHARD_SLICED_VERTEX
	if (const char* e = getenv("B2R_RING_RELAXED")) c->relaxedChain = atoi(e) != 0;
	if (const char* e = getenv("B2R_DIRECT_MAX_AREA")) { const int a = atoi(e); c->directMaxArea = (a >= 0 && a <= 4096) ? a : kDirectMaxAreaDefault; }   // see B2R_OPT_DIRECT_MAX_AREA
	auto bail = [&](int code) { g_createError = c->error; b2r_destroy(c); return code; };
#define CUC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(c, B2R_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); return bail(B2R_E_CUDA); } } while (0)
	CUC(cudaSetDevice(device));
	CUC(cudaDeviceGetAttribute(&c->numSMs, cudaDevAttrMultiProcessorCount, device));
	CUC(cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking));
	c->stream = c->ownStream;
	{
		// B2R_FRONT_PRIORITY=1: the second stream (vertex stage / pipelined front end) gets the highest priority, so its
		// short kernels are dispatched ahead of the remaining CTAs of the long raster grid they overlap with
		int prLeast = 0, prGreatest = 0;
		CUC(cudaDeviceGetStreamPriorityRange(&prLeast, &prGreatest));
		const char* e = getenv("B2R_FRONT_PRIORITY");
		const bool high = e ? atoi(e) != 0 : false;
		CUC(cudaStreamCreateWithPriority(&c->vtxStream, cudaStreamNonBlocking, high ? prGreatest : prLeast));
		CUC(cudaStreamCreateWithPriority(&c->frontStream, cudaStreamNonBlocking, high ? prGreatest : prLeast));
	}
	CUC(cudaEventCreateWithFlags(&c->forkEv, cudaEventDisableTiming));
	for (int i = 0; i < 2; ++i)
	{
		CUC(cudaEventCreateWithFlags(&c->vtxDone[i], cudaEventDisableTiming));
		CUC(cudaEventCreateWithFlags(&c->frontDone[i], cudaEventDisableTiming));
		CUC(cudaEventCreateWithFlags(&c->rasterDone[i], cudaEventDisableTiming));
		CUC(cudaEventCreateWithFlags(&c->listsReady[i], cudaEventDisableTiming));
	}
	for (int i = 0; i < kMaxFrameRing; ++i)
	{
		CUC(cudaStreamCreateWithFlags(&c->ringStream[i], cudaStreamNonBlocking));
		CUC(cudaEventCreateWithFlags(&c->ringDone[i], cudaEventDisableTiming));
		CUC(cudaEventCreateWithFlags(&c->ringRasterDone[i], cudaEventDisableTiming));
	}
	for (auto& slot : c->slots)
	{
		CUC(cudaEventCreateWithFlags(&slot.done, cudaEventDisableTiming));
		for (auto& e : slot.ev) CUC(cudaEventCreate(&e));
		CUC(cudaMallocHost((void**)&slot.hInfo, sizeof(b2r_context::HostInfo)));
		CUC(cudaMallocHost((void**)&slot.hViewProj, 64));
		memset(slot.hInfo, 0, sizeof(b2r_context::HostInfo));
	}
	CUC(cudaFuncSetAttribute(bin_sort_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortLongSmem));
#undef CUC
	int r = allocSurfaces(c);
	if (r) return bail(r);
	r = setBins(c, (width + kTile - 1) / kTile, (height + kTile - 1) / kTile, kTile, kTile);
	if (r) return bail(r);
	*out = c;
	return B2R_OK;
}

void b2r_destroy(b2r_context* c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	if (c->stream) cudaStreamSynchronize(c->stream);
	if (c->vtxStream) { cudaStreamSynchronize(c->vtxStream); cudaStreamDestroy(c->vtxStream); }
	if (c->frontStream) { cudaStreamSynchronize(c->frontStream); cudaStreamDestroy(c->frontStream); }
	if (c->forkEv) cudaEventDestroy(c->forkEv);
	for (int i = 0; i < kMaxFrameRing; ++i)
	{
		if (c->ringStream[i]) { cudaStreamSynchronize(c->ringStream[i]); cudaStreamDestroy(c->ringStream[i]); }
		if (c->ringDone[i]) cudaEventDestroy(c->ringDone[i]);
		if (c->ringRasterDone[i]) cudaEventDestroy(c->ringRasterDone[i]);
	}
	cudaFree(c->dRangeScratch);
	for (int i = 0; i < 2; ++i) { if (c->vtxDone[i]) cudaEventDestroy(c->vtxDone[i]); if (c->frontDone[i]) cudaEventDestroy(c->frontDone[i]); if (c->rasterDone[i]) cudaEventDestroy(c->rasterDone[i]); if (c->listsReady[i]) cudaEventDestroy(c->listsReady[i]); }
	for (auto& t : c->textures) if (t.alive) cudaFree(t.px);
	c->dIVertex.release(); c->dIndices.release(); c->dPosIn.release(); for (auto& v : c->dPos4) v.release(); c->dOvertexScratch.release();
	c->dTriRefs.release(); c->dColourOut.release(); c->dVsChunks.release(); c->dTriChunks.release();
	for (auto& v : c->dMeshConst) v.release(); c->dMeshStatic.release(); c->dTexDesc.release();
	for (auto& B : c->bins)
	{
		B.rects.release(); B.tileCount.release(); B.tileOffset.release(); B.tileCursor.release(); B.longTiles.release();
		B.tileOrder.release(); B.bigList.release(); B.binIds.release(); B.keys.release(); B.dirty.release(); B.binSlots.release();
		cudaFree(B.frameInfo);
	}
	cudaFree(c->dFramebuffer); cudaFree(c->dDepth); cudaFree(c->dZ); cudaFree(c->dWinner);
	releaseShard(c);
	if (c->hostFrame) cudaHostUnregister(c->hostFrame);
	if (c->copyStream) cudaStreamDestroy(c->copyStream);
	if (c->frameDone) cudaEventDestroy(c->frameDone);
	for (auto& e : c->copyDone) if (e) cudaEventDestroy(e);
	for (auto& slot : c->slots)
	{
		if (slot.hStatic) cudaFreeHost(slot.hStatic);
		if (slot.hViewProj) cudaFreeHost(slot.hViewProj);
		if (slot.hInfo) cudaFreeHost(slot.hInfo);
		if (slot.done) cudaEventDestroy(slot.done);
		for (auto& e : slot.ev) if (e) cudaEventDestroy(e);
	}
	if (c->ownStream) cudaStreamDestroy(c->ownStream);
	delete c;
}

int b2r_setup_bins(b2r_context* c, uint32_t bx, uint32_t by, uint32_t bw, uint32_t bh, uint32_t /*queue_max_size*/)
{
	if (!c) return B2R_E_INVALID;
	CU(c, cudaSetDevice(c->device));
	// the shared arena (tile tables, inboxes of every rank) was sized for the grid at b2r_shard_prepare
	if (c->shard.prepared) return fail(c, B2R_E_STATE, "setup_bins: the bin grid is fixed while the sharded front end is prepared (b2r_shard_release first)");
	int r = validateFrame(c); if (r) return r;
	return setBins(c, bx, by, bw, bh);
}

int b2r_texture_create_rgba8(b2r_context* c, const uint32_t* pixels, uint32_t w, uint32_t h, int* outId)
{
	if (!c) return B2R_E_INVALID;
	if (!pixels || !w || !h || !outId) return fail(c, B2R_E_INVALID, "texture arguments invalid");
	CU(c, cudaSetDevice(c->device));
	HostTexture t; t.w = w; t.h = h; t.alive = true;
	CU(c, cudaMalloc((void**)&t.px, (size_t)w * h * 4));
	// round-2 experiment (variant build -DB2R_TEX_TILED_LAYOUT + B2R_TEX_TILED=1; measured: no gain): 8x4-texel tiles
	std::vector<uint32_t> tiled;
#ifdef B2R_TEX_TILED_LAYOUT
	static const bool wantTiled = [] { const char* e = getenv("B2R_TEX_TILED"); return e && atoi(e) != 0; }();
#else
	const bool wantTiled = false;
#endif
	if (wantTiled && w >= 8 && (w & (w - 1)) == 0 && h % 4 == 0)
	{
		while ((1u << t.tiledShift) < w) ++t.tiledShift;
		tiled.resize((size_t)w * h);
		for (uint32_t y = 0; y < h; ++y)
			for (uint32_t x = 0; x < w; ++x)
				tiled[((((size_t)(y >> 2) * (w >> 3) + (x >> 3)) << 5) | ((y & 3u) << 3) | (x & 7u))] = pixels[(size_t)y * w + x];
		pixels = tiled.data();
	}
	cudaError_t e = cudaMemcpyAsync(t.px, pixels, (size_t)w * h * 4, cudaMemcpyHostToDevice, c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	if (e != cudaSuccess) { cudaFree(t.px); return fail(c, B2R_E_CUDA, "texture upload failed: %s", cudaGetErrorString(e)); }
	c->textures.push_back(t);
	c->texDirty = true; c->meshStaticDirty = true;
	*outId = (int)c->textures.size() - 1;
	return B2R_OK;
}

int b2r_texture_destroy(b2r_context* c, int id)
{
	if (!c) return B2R_E_INVALID;
	if (!texValid(c, id)) return fail(c, B2R_E_INVALID, "texture id %d is not alive", id);
	CU(c, cudaSetDevice(c->device));
	int r = validateFrame(c); if (r) return r;
	cudaFree(c->textures[id].px);
	c->textures[id] = HostTexture{};
	c->texDirty = true; c->meshStaticDirty = true;
	return B2R_OK;
}

static int addMesh(b2r_context* c, HostMesh hm, const float world[16], const int tex[4], int* outId)
{
	memcpy(hm.world, world, 64);
	for (int i = 0; i < 4; ++i) hm.texIds[i] = tex ? tex[i] : -1;
	hm.vtxOutBase = (uint32_t)c->nVtxOut; hm.primBase = (uint32_t)c->nPrims;
	if (c->nVtxOut + hm.nVertices > 0xFFFFFFF0ull || c->nPrims + hm.nTris > 0xFFFFFFF0ull)
		return fail(c, B2R_E_INVALID, "scene exceeds 2^32 vertices or primitives");
	c->nVtxOut += hm.nVertices; c->nPrims += hm.nTris;
	c->meshes.push_back(hm);
	c->sceneDirty = true; c->meshStaticDirty = true;
	if (outId) *outId = (int)c->meshes.size() - 1;
	return B2R_OK;
}

int b2r_mesh_upload(b2r_context* c, const void* ivertex56, uint32_t nV, const uint32_t* indices, uint32_t nI, int topology,
                    const float world[16], const int tex[4], int* outId)
{
	if (!c) return B2R_E_INVALID;
	if (topology != B2R_TOPOLOGY_TRIANGLE_LIST)
		return fail(c, B2R_E_UNSUPPORTED, "only TriangleList is supported (TriangleStrip is broken on the reference GPU path, CUDARenderer.cu:733-753)");
	if (!world || (nV && !ivertex56) || (nI && !indices)) return fail(c, B2R_E_INVALID, "null mesh argument");
	for (uint32_t i = 0; i < nI - nI % 3; ++i)
		if (indices[i] >= nV) return fail(c, B2R_E_INVALID, "index %u = %u out of range (%u vertices)", i, indices[i], nV);
	CU(c, cudaSetDevice(c->device));
	int r = validateFrame(c); if (r) return r;
	CU(c, c->dIVertex.reserve((c->nVtxIn + nV) * 7 + 1, c->stream, c->nVtxIn * 7));
	CU(c, c->dIndices.reserve(c->nIdx + nI + 1, c->stream, c->nIdx));
	CU(c, c->dPosIn.reserve(c->nVtxIn + nV + 1, c->stream, c->nVtxIn));
	if (nV)
	{
		// on the context's stream: a plain cudaMemcpy from pageable memory may return before its DMA has landed and
		// is ordered only against the legacy default stream, not against this (non-blocking) stream
		CU(c, cudaMemcpyAsync(c->dIVertex.ptr + c->nVtxIn * 7, ivertex56, (size_t)nV * 56, cudaMemcpyHostToDevice, c->stream));
		split_positions_kernel<<<(nV + 255) / 256, 256, 0, c->stream>>>(c->dIVertex.ptr, c->dPosIn.ptr, (uint32_t)c->nVtxIn, nV);
		CU(c, cudaGetLastError());
	}
	if (nI) CU(c, cudaMemcpyAsync(c->dIndices.ptr + c->nIdx, indices, (size_t)nI * 4, cudaMemcpyHostToDevice, c->stream));
	CU(c, cudaStreamSynchronize(c->stream));       // the caller keeps ownership of the host arrays
	HostMesh hm{};
	hm.vtxInBase = (uint32_t)c->nVtxIn; hm.nVertices = nV; hm.idxBase = (uint32_t)c->nIdx; hm.nTris = nI / 3;
	c->nVtxIn += nV; c->nIdx += nI;
	return addMesh(c, hm, world, tex, outId);
}

int b2r_mesh_instance(b2r_context* c, int src, const float world[16], const int tex[4], int* outId)
{
	if (!c) return B2R_E_INVALID;
	if (src < 0 || (size_t)src >= c->meshes.size() || !world) return fail(c, B2R_E_INVALID, "bad source mesh id %d", src);
	CU(c, cudaSetDevice(c->device));
	int r = validateFrame(c); if (r) return r;
	HostMesh hm = c->meshes[src];
	return addMesh(c, hm, world, tex ? tex : c->meshes[src].texIds, outId);
}

int b2r_mesh_set_world(b2r_context* c, int id, const float world[16])
{
	if (!c) return B2R_E_INVALID;
	if (id < 0 || (size_t)id >= c->meshes.size() || !world) return fail(c, B2R_E_INVALID, "bad mesh id %d", id);
	memcpy(c->meshes[id].world, world, 64);
	c->meshStaticDirty = true;
	return B2R_OK;
}

int b2r_mesh_clear_all(b2r_context* c)
{
	if (!c) return B2R_E_INVALID;
	CU(c, cudaSetDevice(c->device));
	int r = validateFrame(c); if (r) return r;
	c->meshes.clear();
	c->nVtxIn = c->nVtxOut = c->nIdx = c->nPrims = 0;
	c->sceneDirty = true; c->meshStaticDirty = true;
	return B2R_OK;
}

int b2r_set_option(b2r_context* c, int option, int value)
{
	if (!c) return B2R_E_INVALID;
	switch (option)
	{
	case B2R_OPT_FILL_RULE_TOP_LEFT: c->fillTopLeft = value ? 1 : 0; break;
	case B2R_OPT_TILE_OWNER_RANK:
	case B2R_OPT_TILE_OWNER_COUNT:
	case B2R_OPT_PACKED_OUTPUT:
		// the sharded front end fixes ownership (rank, count) and the output layout at b2r_shard_prepare: the raster kernel
		// indexes the per-rank tables of the arena with them
		if (c->shard.prepared) return fail(c, B2R_E_STATE, "tile ownership / packed output cannot change while the sharded front end is prepared");
		if (option == B2R_OPT_TILE_OWNER_RANK) { if (value < 0) return fail(c, B2R_E_INVALID, "rank < 0"); c->ownerRank = value; }
		else if (option == B2R_OPT_TILE_OWNER_COUNT) { if (value < 1) return fail(c, B2R_E_INVALID, "count < 1"); c->ownerCount = value; }
		else c->packedOutput = value ? 1 : 0;
		break;
	case B2R_OPT_DEBUG_SURFACES: c->debugSurfaces = value ? 1 : 0; break;
	case B2R_OPT_STAGE_TIMERS: c->stageTimers = value ? 1 : 0; break;
	case B2R_OPT_SORT_BINS: c->sortBins = value ? 1 : 0; break;
	case B2R_OPT_COUNT_FRAGMENTS: c->countFragments = value ? 1 : 0; break;
	case B2R_OPT_OVERLAP_VERTEX: c->overlapVertex = value ? 1 : 0; break;
	case B2R_OPT_OVERLAP_FRONT: c->overlapFront = value ? 1 : 0; break;
	case B2R_OPT_DIRECT_MAX_AREA:
	{
		if (value < 0 || value > 4096) return fail(c, B2R_E_INVALID, "direct-path area must be 0 (off) .. 4096 pixels");
		// the key buffers exist only while the path is on: between frames
		CU(c, cudaSetDevice(c->device));
		int r = validateFrame(c); if (r) return r;
		c->directMaxArea = value;
		for (auto& B : c->bins)
			if (B.frameInfo) { r = reserveBinSet(c, B); if (r) return r; }
		break;
	}
	case B2R_OPT_SHARD_LOCAL_COLOUR:
	{
		// changes where the frames' colours go: only between frames
		CU(c, cudaSetDevice(c->device));
		int r = validateFrame(c); if (r) return r;
		c->shardLocalColour = value ? 1 : 0;
		if (c->shard.bound) c->shard.waitedFrame = c->shard.frame;
		break;
	}
	case B2R_OPT_FRAME_RING:
	case B2R_OPT_SHARD_SLICED_VERTEX:
	{
		// both change which buffers / streams a frame uses: only between frames, and not while the shared arena exists
		if (option == B2R_OPT_FRAME_RING && value != 0 && (value < 2 || value > kMaxFrameRing))
			return fail(c, B2R_E_INVALID, "frame ring depth must be 0 (off) or 2..%d", kMaxFrameRing);
		if (c->shard.prepared) return fail(c, B2R_E_STATE, "set this option before b2r_shard_prepare");
		CU(c, cudaSetDevice(c->device));
		int r = validateFrame(c); if (r) return r;
		if (option == B2R_OPT_FRAME_RING) { c->ringDepth = value; c->sceneDirty = true; }
		else c->slicedVertex = value ? 1 : 0;
		break;
	}
	default: return fail(c, B2R_E_INVALID, "unknown option %d", option);
	}
	if (c->ownerRank >= c->ownerCount) { /* tolerated while the pair is being set; checked at render */ }
	return B2R_OK;
}

int b2r_get_option(b2r_context* c, int option, int* value)
{
	if (!c) return B2R_E_INVALID;
	if (!value) return fail(c, B2R_E_INVALID, "null output");
	switch (option)
	{
	case B2R_OPT_FILL_RULE_TOP_LEFT: *value = c->fillTopLeft; break;
	case B2R_OPT_TILE_OWNER_RANK: *value = c->ownerRank; break;
	case B2R_OPT_TILE_OWNER_COUNT: *value = c->ownerCount; break;
	case B2R_OPT_DEBUG_SURFACES: *value = c->debugSurfaces; break;
	case B2R_OPT_STAGE_TIMERS: *value = c->stageTimers; break;
	case B2R_OPT_SORT_BINS: *value = c->sortBins; break;
	case B2R_OPT_COUNT_FRAGMENTS: *value = c->countFragments; break;
	case B2R_OPT_PACKED_OUTPUT: *value = c->packedOutput; break;
	case B2R_OPT_OVERLAP_VERTEX: *value = c->overlapVertex; break;
	case B2R_OPT_OVERLAP_FRONT: *value = c->overlapFront; break;
	case B2R_OPT_DIRECT_MAX_AREA: *value = c->directMaxArea; break;
	case B2R_OPT_FRAME_RING: *value = c->ringDepth; break;
	case B2R_OPT_SHARD_SLICED_VERTEX: *value = c->slicedVertex; break;
	case B2R_OPT_SHARD_LOCAL_COLOUR: *value = c->shardLocalColour; break;
	default: return fail(c, B2R_E_INVALID, "unknown option %d", option);
	}
	return B2R_OK;
}

int b2r_render(b2r_context* c, const float camPos[3], const float camFwd[3], const float viewProj[16],
               int cull, int sample, int vis, const float clearRGB[3])
{
	if (!c) return B2R_E_INVALID;
	if (!camPos || !camFwd || !viewProj) return fail(c, B2R_E_INVALID, "null camera argument");
	if (cull < 0 || cull > 2 || sample < 0 || sample > 1 || vis < 0 || vis > 2) return fail(c, B2R_E_INVALID, "bad render state");
	if (c->ownerRank >= c->ownerCount) return fail(c, B2R_E_INVALID, "tile owner rank %d >= count %d", c->ownerRank, c->ownerCount);
	if (c->packedOutput && !c->boundFramebuffer)
	{
		// tile-major output holds whole 16x16 blocks: more pixels than the W x H surface when the screen is ragged
		const uint64_t subs = (uint64_t)((c->binW + kTile - 1) / kTile) * ((c->binH + kTile - 1) / kTile);
		const uint64_t tiles = nBins(c), owned = tiles > (uint64_t)c->ownerRank ? (tiles - c->ownerRank + c->ownerCount - 1) / c->ownerCount : 0;
		if (owned * subs * kTile * kTile > (uint64_t)c->width * c->height)
			return fail(c, B2R_E_STATE, "packed output of %llu pixels does not fit the %ux%u colour surface: bind a buffer of b2r_packed_size pixels (b2r_bind_framebuffer)",
				(unsigned long long)(owned * subs * kTile * kTile), c->width, c->height);
	}
	CU(c, cudaSetDevice(c->device));
	RenderCall& rc = c->last;
	memcpy(rc.camPos, camPos, 12); memcpy(rc.camFwd, camFwd, 12); memcpy(rc.viewProj, viewProj, 64);
	static const float defClear[3] = { 0.25f, 0.25f, 0.25f };   // CUDARenderer.h:135
	memcpy(rc.clearRGB, clearRGB ? clearRGB : defClear, 12);
	rc.cull = cull; rc.sample = sample; rc.vis = vis; rc.valid = true;
	return enqueueFrame(c);
}

int b2r_sync(b2r_context* c)
{
	if (!c) return B2R_E_INVALID;
	CU(c, cudaSetDevice(c->device));
	return validateFrame(c);
}

static int readSurface(b2r_context* c, const void* src, void* out, const char* what)
{
	if (!out) return fail(c, B2R_E_INVALID, "null output");
	if (!c->haveFrame) return fail(c, B2R_E_STATE, "%s: nothing rendered yet", what);
	if (!src) return fail(c, B2R_E_STATE, "%s: surface not enabled (B2R_OPT_DEBUG_SURFACES)", what);
	CU(c, cudaSetDevice(c->device));
	int r = validateFrame(c); if (r) return r;
	CU(c, cudaMemcpyAsync(out, src, (size_t)c->width * c->height * 4, cudaMemcpyDeviceToHost, c->stream));
	CU(c, cudaStreamSynchronize(c->stream));
	return B2R_OK;
}

int b2r_read_framebuffer(b2r_context* c, uint32_t* out)
{
	if (!c) return B2R_E_INVALID;
	return readSurface(c, c->boundFramebuffer ? c->boundFramebuffer : c->dColour[c->lastColourIdx], out, "read_framebuffer");
}
int b2r_draw_texture(b2r_context* c, const uint32_t* pixels, uint32_t w, uint32_t h, int global, int stretched, int sample, const float clearRGB[3])
{
	if (!c) return B2R_E_INVALID;
	if (!pixels || !w || !h || sample < 0 || sample > 1) return fail(c, B2R_E_INVALID, "draw_texture: bad arguments");
	if (c->boundFramebuffer) return fail(c, B2R_E_STATE, "draw_texture: the colour surface is bound to caller memory");
	CU(c, cudaSetDevice(c->device));
	int r = validateFrame(c); if (r) return r;
	uint32_t* d = nullptr;
	CU(c, cudaMalloc((void**)&d, (size_t)w * h * 4));
	cudaError_t e = cudaMemcpyAsync(d, pixels, (size_t)w * h * 4, cudaMemcpyHostToDevice, c->stream);
	if (e == cudaSuccess)
	{
		static const float defClear[3] = { 0.25f, 0.25f, 0.25f };
		const float* cl = clearRGB ? clearRGB : defClear;
		const uint32_t clear = ((uint32_t)(cl[2] * 255.f) & 0xFFu) | (((uint32_t)(cl[1] * 255.f) & 0xFFu) << 8) | (((uint32_t)(cl[0] * 255.f) & 0xFFu) << 16);
		draw_texture_kernel<<<dim3((c->width + 15) / 16, (c->height + 15) / 16), 256, 0, c->stream>>>(c->dColour[c->colourIdx], TexDesc{ d, w, h, 0u, 0u }, c->width, c->height,
			global ? 1u : 0u, stretched ? 1u : 0u, (uint32_t)sample, clear);
		e = cudaGetLastError();
	}
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	cudaFree(d);
	if (e != cudaSuccess) return fail(c, B2R_E_CUDA, "draw_texture failed: %s", cudaGetErrorString(e));
	c->lastColourIdx = c->colourIdx;
	c->haveFrame = true;
	return B2R_OK;
}

static int ensureCopyStream(b2r_context* c)
{
	if (!c->copyStream)
	{
		CU(c, cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
		CU(c, cudaEventCreateWithFlags(&c->frameDone, cudaEventDisableTiming));
		for (auto& e : c->copyDone) CU(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	}
	return B2R_OK;
}

int b2r_present_async(b2r_context* c, uint32_t* hostPixels)
{
	if (!c) return B2R_E_INVALID;
	if (!hostPixels) return fail(c, B2R_E_INVALID, "null output");
	if (!c->haveFrame) return fail(c, B2R_E_STATE, "present_async: nothing rendered yet");
	if (c->boundFramebuffer) return fail(c, B2R_E_STATE, "present_async: the colour surface is bound to caller memory");
	CU(c, cudaSetDevice(c->device));
	const size_t bytes = (size_t)c->width * c->height * 4;
	{ int r = ensureCopyStream(c); if (r) return r; }
	const int idx = c->lastColourIdx;
	{ int r = enqueueColourWait(c); if (r) return r; }
	CU(c, cudaEventRecord(c->frameDone, c->stream));
	CU(c, cudaStreamWaitEvent(c->copyStream, c->frameDone, 0));
	CU(c, cudaMemcpyAsync(hostPixels, c->dColour[idx], bytes, cudaMemcpyDeviceToHost, c->copyStream));
	CU(c, cudaEventRecord(c->copyDone[idx], c->copyStream));
	c->copyPending[idx] = true;
	if (!c->shard.bound) c->colourIdx = idx ^ 1;          // the next frame renders into the other surface (sharded: frame parity)
	return B2R_OK;
}

int b2r_host_frame_register(b2r_context* c, void* host, uint64_t bytes)
{
	if (!c) return B2R_E_INVALID;
	CU(c, cudaSetDevice(c->device));
	if (c->copyStream) CU(c, cudaStreamSynchronize(c->copyStream));
	int r = validateFrame(c); if (r) return r;
	if (c->hostFrame)
	{
		CU(c, cudaHostUnregister(c->hostFrame));
		c->hostFrame = nullptr; c->hostFrameDev = nullptr; c->hostFrameBytes = 0;
	}
	if (!host && !bytes) return B2R_OK;
	if (!host || bytes < (uint64_t)c->width * c->height * 4) return fail(c, B2R_E_INVALID, "host frame: need at least width*height*4 bytes");
	CU(c, cudaHostRegister(host, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
	void* dev = nullptr;
	cudaError_t e = cudaHostGetDevicePointer(&dev, host, 0);
	if (e != cudaSuccess) { cudaHostUnregister(host); return fail(c, B2R_E_CUDA, "cudaHostGetDevicePointer failed: %s", cudaGetErrorString(e)); }
	c->hostFrame = host; c->hostFrameDev = (uint32_t*)dev; c->hostFrameBytes = bytes;
	return B2R_OK;
}

int b2r_present_tiles_async(b2r_context* c, uint64_t byteOffset)
{
	if (!c) return B2R_E_INVALID;
	if (!c->haveFrame) return fail(c, B2R_E_STATE, "present_tiles_async: nothing rendered yet");
	if (!c->hostFrameDev) return fail(c, B2R_E_STATE, "present_tiles_async: no host frame registered (b2r_host_frame_register)");
	if (c->boundFramebuffer || c->packedOutput) return fail(c, B2R_E_STATE, "present_tiles_async: needs the context's own row-major colour surface");
	if (c->shard.bound && !c->shardLocalColour) return fail(c, B2R_E_STATE, "present_tiles_async: with the sharded front end set B2R_OPT_SHARD_LOCAL_COLOUR");
	const uint64_t frameBytes = (uint64_t)c->width * c->height * 4;
	if ((byteOffset & 15u) || byteOffset + frameBytes > c->hostFrameBytes) return fail(c, B2R_E_INVALID, "present_tiles_async: offset %llu outside the registered host frame", (unsigned long long)byteOffset);
	CU(c, cudaSetDevice(c->device));
	{ int r = ensureCopyStream(c); if (r) return r; }
	const int idx = c->lastColourIdx;
	const FrameParams fp = makeParams(c, c->last);
	const uint32_t tiles = nBins(c);
	const uint32_t owned = (tiles > (uint32_t)c->ownerRank) ? (tiles - (uint32_t)c->ownerRank + (uint32_t)c->ownerCount - 1) / (uint32_t)c->ownerCount : 0;
	CU(c, cudaEventRecord(c->frameDone, c->stream));
	CU(c, cudaStreamWaitEvent(c->copyStream, c->frameDone, 0));
	if (owned)
	{
		const uint32_t blocks = owned * fp.subX * fp.subY;
		present_tiles_kernel<<<(blocks + 3) / 4, 256, 0, c->copyStream>>>(c->dColour[idx], c->hostFrameDev + byteOffset / 4, (uint32_t)c->ownerRank, (uint32_t)c->ownerCount, owned, fp);
		CU(c, cudaGetLastError());
	}
	CU(c, cudaEventRecord(c->copyDone[idx], c->copyStream));
	c->copyPending[idx] = true;
	if (!c->shard.bound) c->colourIdx = idx ^ 1;          // the next frame renders into the other surface (sharded: frame parity)
	return B2R_OK;
}

static int presentRows(b2r_context* c, uint64_t byteOffset, bool explicitRows, uint32_t rowBegin, uint32_t rowEnd)
{
	if (!c) return B2R_E_INVALID;
	if (!c->haveFrame) return fail(c, B2R_E_STATE, "present_band_async: nothing rendered yet");
	if (!c->hostFrame) return fail(c, B2R_E_STATE, "present_band_async: no host frame registered (b2r_host_frame_register)");
	if (c->boundFramebuffer || c->packedOutput) return fail(c, B2R_E_STATE, "present_band_async: needs the row-major colour surface");
	if (c->shard.bound && c->shardLocalColour) return fail(c, B2R_E_STATE, "present_band_async: the frame is not composed on rank 0 (B2R_OPT_SHARD_LOCAL_COLOUR is set)");
	const uint64_t frameBytes = (uint64_t)c->width * c->height * 4;
	if ((byteOffset & 15u) || byteOffset + frameBytes > c->hostFrameBytes) return fail(c, B2R_E_INVALID, "present_band_async: offset %llu outside the registered host frame", (unsigned long long)byteOffset);
	CU(c, cudaSetDevice(c->device));
	{ int r = ensureCopyStream(c); if (r) return r; }
	b2r_context::Shard& sh = c->shard;
	const int idx = c->lastColourIdx;
	// this rank's band: whole rows of bins
	const uint32_t n = sh.bound ? (uint32_t)sh.n : 1u, self = sh.bound ? (uint32_t)sh.self : 0u;
	const uint32_t binRows = (c->binsY + n - 1) / n;
	uint64_t row0 = std::min<uint64_t>((uint64_t)self * binRows * c->binH, c->height), row1 = std::min<uint64_t>((uint64_t)(self + 1) * binRows * c->binH, c->height);
	if (explicitRows)
	{
		if (rowBegin > rowEnd || rowEnd > c->height) return fail(c, B2R_E_INVALID, "present_rows_async: rows [%u, %u) outside the %u rows of the frame", rowBegin, rowEnd, c->height);
		row0 = rowBegin; row1 = rowEnd;
	}
	CU(c, cudaEventRecord(c->frameDone, c->stream));
	CU(c, cudaStreamWaitEvent(c->copyStream, c->frameDone, 0));
	const uint32_t* src = c->dColour[idx];
	if (sh.bound)
	{
		// every rank's colours of the frame have landed in rank 0's surface (flag words 8..15 of rank 0's arena, polled
		// over NVLink by the other ranks)
		PeerFlags pf{};
		for (int p = 0; p < sh.n; ++p) pf.flags[p] = reinterpret_cast<uint32_t*>(sh.self == 0 ? sh.arena : sh.peerArena[0]);
		CU(c, launchFrameKernel(c, peer_barrier_kernel, dim3(1), dim3(32), 0, c->copyStream, pf, (uint32_t)sh.self, (uint32_t)sh.n, 8u, sh.frame, 0u, 1u, sh.hTimeout, 1u));
		src = sh.colourBase + (size_t)idx * c->width * c->height;
	}
	if (row1 > row0)
		CU(c, cudaMemcpyAsync(static_cast<uint8_t*>(c->hostFrame) + byteOffset + row0 * c->width * 4, src + row0 * c->width, (row1 - row0) * c->width * 4, cudaMemcpyDeviceToHost, c->copyStream));
	CU(c, cudaEventRecord(c->copyDone[idx], c->copyStream));
	c->copyPending[idx] = true;
	if (!sh.bound) c->colourIdx = idx ^ 1;          // the next frame renders into the other surface (sharded: frame parity)
	return B2R_OK;
}

int b2r_present_band_async(b2r_context* c, uint64_t byteOffset) { return presentRows(c, byteOffset, false, 0u, 0u); }
int b2r_present_rows_async(b2r_context* c, uint64_t byteOffset, uint32_t rowBegin, uint32_t rowEnd) { return presentRows(c, byteOffset, true, rowBegin, rowEnd); }

int b2r_present_wait(b2r_context* c)
{
	if (!c) return B2R_E_INVALID;
	CU(c, cudaSetDevice(c->device));
	if (c->copyStream) CU(c, cudaStreamSynchronize(c->copyStream));
	c->copyPending[0] = c->copyPending[1] = false;
	return validateFrame(c);
}

int b2r_read_depth(b2r_context* c, int32_t* out) { if (!c) return B2R_E_INVALID; return readSurface(c, c->dDepthOf[c->lastColourIdx], out, "read_depth"); }
int b2r_read_z(b2r_context* c, float* out) { if (!c) return B2R_E_INVALID; return readSurface(c, c->dZ, out, "read_z"); }
int b2r_read_winner(b2r_context* c, uint32_t* out) { if (!c) return B2R_E_INVALID; return readSurface(c, c->dWinner, out, "read_winner"); }

int b2r_read_overtex(b2r_context* c, int id, float* out18)
{
	if (!c) return B2R_E_INVALID;
	if (id < 0 || (size_t)id >= c->meshes.size() || !out18) return fail(c, B2R_E_INVALID, "bad mesh id %d", id);
	if (!c->haveFrame) return fail(c, B2R_E_STATE, "read_overtex: nothing rendered yet");
	CU(c, cudaSetDevice(c->device));
	int r = validateFrame(c); if (r) return r;
	const HostMesh& hm = c->meshes[id];
	if (!hm.nVertices) return B2R_OK;
	if (c->ringDepth && c->shard.bound && c->slicedVertex)
		return fail(c, B2R_E_STATE, "read_overtex: with the sliced vertex stage a rank holds only the transformed vertices of its own primitive slice");
	CU(c, c->dOvertexScratch.reserve((size_t)hm.nVertices * 18, c->stream));
	const FrameParams fp = makeParams(c, c->last);
	overtex_kernel<<<(hm.nVertices + 255) / 256, 256, 0, c->stream>>>(c->dMeshConst[c->geoIdx].ptr, (uint32_t)id, c->dIVertex.ptr, c->dPos4[c->geoIdx].ptr, c->dOvertexScratch.ptr, fp);
	CU(c, cudaGetLastError());
	CU(c, cudaMemcpyAsync(out18, c->dOvertexScratch.ptr, (size_t)hm.nVertices * 18 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
	CU(c, cudaStreamSynchronize(c->stream));
	return B2R_OK;
}

int b2r_read_visible(b2r_context* c, int id, uint8_t* out)
{
	if (!c) return B2R_E_INVALID;
	if (id < 0 || (size_t)id >= c->meshes.size() || !out) return fail(c, B2R_E_INVALID, "bad mesh id %d", id);
	if (!c->haveFrame) return fail(c, B2R_E_STATE, "read_visible: nothing rendered yet");
	CU(c, cudaSetDevice(c->device));
	int r = validateFrame(c); if (r) return r;
	const HostMesh& hm = c->meshes[id];
	std::vector<uint2> rc(hm.nTris);
	if (hm.nTris) CU(c, cudaMemcpy(rc.data(), c->curRects + hm.primBase, (size_t)hm.nTris * 8, cudaMemcpyDeviceToHost));
	for (uint32_t t = 0; t < hm.nTris; ++t) out[t] = ((rc[t].x & 0xFFFFu) == 0xFFFFu) ? 0 : 1;
	return B2R_OK;
}

int b2r_read_bins(b2r_context* c, uint32_t* counts, uint32_t* offsets, uint32_t* ids, uint64_t idsCap, uint64_t* total)
{
	if (!c) return B2R_E_INVALID;
	if (!c->haveFrame) return fail(c, B2R_E_STATE, "read_bins: nothing rendered yet");
	CU(c, cudaSetDevice(c->device));
	int r = validateFrame(c); if (r) return r;
	if (c->lastDirect && c->shard.bound)
		return fail(c, B2R_E_STATE, "read_bins: the newest frame of the sharded front end resolved its small triangles through the key buffers, so this rank's lists "
			"hold the other triangles only; render the frame with B2R_OPT_DIRECT_MAX_AREA = 0 on every rank to read complete lists");
	const bool rerun = c->lastDirect;
	const int savedDirect = c->directMaxArea;
	if (rerun)
	{
		// the newest frame resolved its small triangles in the setup kernel, so its lists hold the other triangles only: the
		// bins of the frame (every visible triangle in every tile of its rectangle, as the reference's TB_Kernel fills
		// them) are produced by running the frame once more with every triangle binned
		c->directMaxArea = 0;
		r = enqueueFrame(c);
		if (!r) r = validateFrame(c);
		c->directMaxArea = savedDirect;
		if (r) return r;
	}
	// (and once the bins have been copied out, the frame is run a third time the way it was rendered, so that the surfaces
	// and counters read afterwards are again those of the frame as the caller rendered it)
	struct Restore
	{
		b2r_context* c; bool on;
		~Restore() { if (on && enqueueFrame(c) == B2R_OK) validateFrame(c); }
	} restore{ c, rerun };
	const uint64_t pairs = c->lastInfo.frameInfo[0];
	if (total) *total = pairs;
	const uint32_t n = nBins(c);
	const bool longLeft = c->sortBins && c->lastLongSkipped && c->lastInfo.frameInfo[3] != 0;      // the frame ran without the long-list sorter but had long lists
	if (ids && (!c->curSorted || longLeft) && c->curSet && c->nPrims && !c->lastInfo.frameInfo[1])
	{
		// B2R_OPT_SORT_BINS = 0: the frame left the lists in scatter order (the image does not depend on it);
		// whoever looks at them gets the canonical order - sort the newest frame's lists now, in place
		r = enqueueListSort(c, *c->curSet, c->curTileCount, c->curTileOffset, c->curBinIds, c->stream, !c->curSorted); if (r) return r;
		CU(c, cudaStreamSynchronize(c->stream));
		// counters of the sorter (tiles beyond its reach, fallbacks) as after an in-frame sort
		CU(c, cudaMemcpy(c->lastInfo.frameInfo, c->curSet->frameInfo, sizeof(c->lastInfo.frameInfo), cudaMemcpyDeviceToHost));
		c->curSorted = true; c->lastLongSkipped = false;
	}
	if (counts) CU(c, cudaMemcpy(counts, c->curTileCount, (size_t)n * 4, cudaMemcpyDeviceToHost));
	if (offsets) CU(c, cudaMemcpy(offsets, c->curTileOffset, (size_t)n * 4, cudaMemcpyDeviceToHost));
	if (ids)
	{
		if (idsCap < pairs) return fail(c, B2R_E_INVALID, "ids capacity %llu < %llu pairs", (unsigned long long)idsCap, (unsigned long long)pairs);
		if (pairs) CU(c, cudaMemcpy(ids, c->curBinIds, (size_t)pairs * 4, cudaMemcpyDeviceToHost));
	}
	return B2R_OK;
}

int b2r_get_stats(b2r_context* c, uint32_t* totalTris, uint32_t* visibleTris)
{
	if (!c) return B2R_E_INVALID;
	CU(c, cudaSetDevice(c->device));
	int r = validateFrame(c); if (r) return r;
	if (totalTris) *totalTris = (uint32_t)c->nPrims;
	if (visibleTris) *visibleTris = c->haveFrame ? c->lastInfo.counters[0] : 0;
	return B2R_OK;
}

int b2r_get_stage_ms(b2r_context* c, float* out, int n)
{
	if (!c) return B2R_E_INVALID;
	if (!out || n < 1) return fail(c, B2R_E_INVALID, "bad output");
	CU(c, cudaSetDevice(c->device));
	int r = validateFrame(c); if (r) return r;
	if (!c->haveStageTimes) return fail(c, B2R_E_STATE, "no frame was rendered with B2R_OPT_STAGE_TIMERS enabled");
	for (int i = 0; i < n && i < B2R_NUM_STAGES; ++i) out[i] = c->stageLast[i];
	return B2R_OK;
}

int b2r_get_stage_ms_mean(b2r_context* c, float* out, int n, uint64_t* frames, int reset)
{
	if (!c) return B2R_E_INVALID;
	if (!out || n < 1) return fail(c, B2R_E_INVALID, "bad output");
	CU(c, cudaSetDevice(c->device));
	int r = validateFrame(c); if (r) return r;
	for (int i = 0; i < n && i < B2R_NUM_STAGES; ++i) out[i] = c->stageFrames ? (float)(c->stageAccum[i] / (double)c->stageFrames) : 0.f;
	if (frames) *frames = c->stageFrames;
	if (reset) { for (auto& a : c->stageAccum) a = 0.0; c->stageFrames = 0; }
	return B2R_OK;
}

int b2r_get_counters(b2r_context* c, uint64_t* out, int n)
{
	if (!c) return B2R_E_INVALID;
	if (!out || n < 1) return fail(c, B2R_E_INVALID, "bad output");
	CU(c, cudaSetDevice(c->device));
	int r = validateFrame(c); if (r) return r;
	const uint32_t tiles = nBins(c);
	const uint64_t owned = (tiles > (uint32_t)c->ownerRank) ? (tiles - c->ownerRank + c->ownerCount - 1) / c->ownerCount : 0;
	const uint64_t v[8] = { c->lastInfo.frameInfo[0] + c->lastInfo.frameInfo[kInfoDirectPairs], c->kernelsLaunched, owned, tiles, c->lastInfo.frameInfo[2], c->lastInfo.counters[1], c->lastInfo.frameInfo[4], c->lastInfo.frameInfo[6] };
	for (int i = 0; i < n && i < 8; ++i) out[i] = v[i];
	return B2R_OK;
}

int b2r_set_stream(b2r_context* c, void* stream)
{
	if (!c) return B2R_E_INVALID;
	CU(c, cudaSetDevice(c->device));
	int r = validateFrame(c); if (r) return r;
	c->stream = stream ? (cudaStream_t)stream : c->ownStream;
	return B2R_OK;
}

int b2r_device_surfaces(b2r_context* c, void** framebuffer, void** depth)
{
	if (!c) return B2R_E_INVALID;
	if (framebuffer) *framebuffer = c->boundFramebuffer ? c->boundFramebuffer : c->dColour[c->lastColourIdx];
	if (depth) *depth = c->dDepthOf[c->lastColourIdx];
	return B2R_OK;
}

static uint32_t ownedTiles(uint32_t tiles, uint32_t rank, uint32_t count)
{
	return tiles > rank ? (tiles - rank + count - 1) / count : 0;
}

int b2r_packed_size(b2r_context* c, int rank, int count, uint64_t* pixels)
{
	if (!c) return B2R_E_INVALID;
	if (!pixels || count < 1 || rank < 0 || rank >= count) return fail(c, B2R_E_INVALID, "bad rank/count");
	const FrameParams fp = makeParams(c, c->last);
	*pixels = (uint64_t)ownedTiles(nBins(c), (uint32_t)rank, (uint32_t)count) * fp.subX * fp.subY * kTile * kTile;
	return B2R_OK;
}

int b2r_unpack_tiles(b2r_context* c, const void* packed, int rank, int count, void* dstFramebuffer)
{
	if (!c) return B2R_E_INVALID;
	if (!packed || count < 1 || rank < 0 || rank >= count) return fail(c, B2R_E_INVALID, "bad unpack arguments");
	CU(c, cudaSetDevice(c->device));
	const FrameParams fp = makeParams(c, c->last);
	const uint32_t owned = ownedTiles(nBins(c), (uint32_t)rank, (uint32_t)count);
	if (!owned) return B2R_OK;
	uint32_t* dst = dstFramebuffer ? (uint32_t*)dstFramebuffer : c->dColour[c->lastColourIdx];
	unpack_tiles_kernel<<<dim3(owned, fp.subX * fp.subY), kTile * kTile, 0, c->stream>>>((const uint32_t*)packed, dst, (uint32_t)rank, (uint32_t)count, fp);
	CU(c, cudaGetLastError());
	return B2R_OK;
}

int b2r_ipc_export(b2r_context* c, void* devicePtr, void* handle64)
{
	if (!c) return B2R_E_INVALID;
	if (!handle64) return fail(c, B2R_E_INVALID, "null handle");
	CU(c, cudaSetDevice(c->device));
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
	CU(c, cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64, devicePtr ? devicePtr : (void*)c->dFramebuffer));
	return B2R_OK;
}

int b2r_ipc_open(b2r_context* c, const void* handle64, void** devicePtr)
{
	if (!c) return B2R_E_INVALID;
	if (!handle64 || !devicePtr) return fail(c, B2R_E_INVALID, "null handle");
	CU(c, cudaSetDevice(c->device));
	cudaIpcMemHandle_t h; memcpy(&h, handle64, 64);
	CU(c, cudaIpcOpenMemHandle(devicePtr, h, cudaIpcMemLazyEnablePeerAccess));
	return B2R_OK;
}

int b2r_ipc_close(b2r_context* c, void* devicePtr)
{
	if (!c) return B2R_E_INVALID;
	CU(c, cudaSetDevice(c->device));
	int r = validateFrame(c); if (r) return r;
	CU(c, cudaIpcCloseMemHandle(devicePtr));
	return B2R_OK;
}

int b2r_shard_prepare(b2r_context* c, int rank, int count, uint64_t listCapacity, void* arenaHandle64, void* colourHandle64)
{
	if (!c) return B2R_E_INVALID;
	if (count < 2 || count > kMaxRanks || rank < 0 || rank >= count || !arenaHandle64 || !colourHandle64)
		return fail(c, B2R_E_INVALID, "shard_prepare: need 2..%d ranks, 0 <= rank < count and two handle buffers", kMaxRanks);
	CU(c, cudaSetDevice(c->device));
	int r = validateFrame(c); if (r) return r;
	if (c->shard.prepared) return fail(c, B2R_E_STATE, "shard_prepare: already prepared (b2r_shard_release first)");
	if (c->boundFramebuffer || c->packedOutput) return fail(c, B2R_E_STATE, "shard_prepare: incompatible with a bound framebuffer / packed output");
	if (c->sceneDirty) { r = rebuildScene(c); if (r) return r; }
	b2r_context::Shard& sh = c->shard;
	sh.n = count; sh.self = rank;
	const size_t tiles = nBins(c);
	sh.tilesPad = (tiles + 4 + 63) & ~(size_t)63;
	const size_t def = std::max<size_t>((size_t)1 << 22, (c->nPrims + count - 1) / count * 4);
	sh.listCap = ((listCapacity ? (size_t)listCapacity : def) + 63) & ~(size_t)63;
	sh.setBytes = ((size_t)(2 + 2 * count) * sh.tilesPad + sh.listCap * (size_t)(1 + count)) * 4;
	// direct path across ranks: the peers push 16x16 key blocks, so the tiles must BE the key buffer's blocks
	sh.direct = c->directMaxArea > 0 && c->binW == (uint32_t)kTile && c->binH == (uint32_t)kTile && c->binsX == (c->width + kTile - 1) / kTile;
	sh.ownedMax = (tiles + count - 1) / count;
	if (sh.direct) sh.setBytes += (size_t)count * (sh.ownedMax * 256 * 8 + sh.tilesPad);
	sh.setBytes = (sh.setBytes + 255) & ~(size_t)255;
	sh.nSets = c->ringDepth ? c->ringDepth : 2;
	const size_t bytes = kArenaFlagBytes + (size_t)sh.nSets * sh.setBytes;
	CU(c, cudaMalloc((void**)&sh.arena, bytes));
	CU(c, cudaMemset(sh.arena, 0, bytes));
	CU(c, finishFills(c));
	CU(c, cudaMallocHost((void**)&sh.hTimeout, 8));
	*sh.hTimeout = 0;
	memset(arenaHandle64, 0, 64); memset(colourHandle64, 0, 64);
	CU(c, cudaIpcGetMemHandle((cudaIpcMemHandle_t*)arenaHandle64, sh.arena));
	if (rank == 0) CU(c, cudaIpcGetMemHandle((cudaIpcMemHandle_t*)colourHandle64, c->dFramebuffer));
	c->ownerRank = rank; c->ownerCount = count;
	sh.prepared = true;
	return B2R_OK;
}

int b2r_shard_bind(b2r_context* c, const void* arenaHandles, const void* colourHandleRank0)
{
	if (!c) return B2R_E_INVALID;
	b2r_context::Shard& sh = c->shard;
	if (!sh.prepared || sh.bound) return fail(c, B2R_E_STATE, "shard_bind: call b2r_shard_prepare first (once)");
	if (!arenaHandles || !colourHandleRank0) return fail(c, B2R_E_INVALID, "shard_bind: null handles");
	CU(c, cudaSetDevice(c->device));
	for (int p = 0; p < sh.n; ++p)
	{
		if (p == sh.self) continue;
		cudaIpcMemHandle_t h; memcpy(&h, (const uint8_t*)arenaHandles + (size_t)p * 64, 64);
		CU(c, cudaIpcOpenMemHandle((void**)&sh.peerArena[p], h, cudaIpcMemLazyEnablePeerAccess));
	}
	if (sh.self == 0) { sh.colourBase = c->dFramebuffer; sh.colourIsPeer = false; }
	else
	{
		cudaIpcMemHandle_t h; memcpy(&h, colourHandleRank0, 64);
		CU(c, cudaIpcOpenMemHandle((void**)&sh.colourBase, h, cudaIpcMemLazyEnablePeerAccess));
		sh.colourIsPeer = true;
	}
	sh.frame = 0;
	sh.bound = true;
	return B2R_OK;
}

int b2r_shard_release(b2r_context* c)
{
	if (!c) return B2R_E_INVALID;
	CU(c, cudaSetDevice(c->device));
	int r = validateFrame(c);
	const bool was = c->shard.prepared;
	releaseShard(c);
	c->colourIdx = 0;
	if (was) { c->ownerRank = 0; c->ownerCount = 1; c->shardLocalColour = 0; }      // b2r_shard_prepare set them: later frames are whole frames again
	return r;
}

int b2r_bind_framebuffer(b2r_context* c, void* devicePtr)
{
	if (!c) return B2R_E_INVALID;
	// takes effect for frames enqueued from now on; frames already in flight keep their target
	c->boundFramebuffer = (uint32_t*)devicePtr;
	return B2R_OK;
}

}
