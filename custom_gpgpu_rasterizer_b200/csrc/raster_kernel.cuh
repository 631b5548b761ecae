// Per-tile raster + deferred shade + ROP write-back in ONE kernel.
// Replaces RA_PerTileKernel, RasterizePixelAtomic, PerformDepthTestAtomic, the 152 B/pixel
// fragment buffer, the three CLEAR_* kernels and PS_Kernel of the reference
// (CUDARenderer.cu:598-643, 833-897, 1039-1129; CUDAROPs.cu:231-251, 329-470).
//
// One CTA = one 16x16 pixel tile, one thread = one pixel.  The tile's depth/winner state
// never leaves the SM: a 64-bit key (z_int << 32 | primitive id) per pixel lives in a
// register for triangles that are rasterised pixel-parallel and in shared memory for small
// triangles that are rasterised triangle-parallel (atomicMin).  Because the resolve is a
// min over (depth, primitive id) it is independent of the order in which triangles are
// visited: equal depths go to the lower primitive id, i.e. to the first-submitted triangle
// (the sequential meaning of the reference's "first to take the lock wins").
// After the last triangle each pixel is shaded once (like the reference's deferred
// PS_Kernel) and colour + depth are written with one coalesced store per row segment.
#pragma once
#include "device_types.cuh"
#include "geometry_kernels.cuh"

namespace b2r {

struct RasterArgs
{
	const uint32_t* tileCount;
	const uint32_t* tileOffset;
	const uint32_t* binIds;     // (the kernel writes the lists it sorts back through these pointers)
	const uint4* triRefs;       // static {v0, v1, v2, mesh} per primitive (global transformed-vertex indices)
	const float4* pos4;
	const float4* posIn;        // model-space positions (x, y, z, 1) of the IVertex records: the OTF instantiation transforms what it needs itself
	const float2* ivertex;      // resident IVertex records (7 float2 each)
	const float4* colour;       // the records' vertex colours, indexed like pos4 (build_colour_stream_kernel)
	const MeshConst* meshes;
	const TexDesc* textures;
	unsigned long long* frameInfo;
	uint32_t* framebuffer;
	int32_t* depth;
	float* zSurface;        // debug surfaces (may be null)
	uint32_t* winnerSurface;
	ShardTable shard;       // sharded front end only (fp.shardGeometry): the per-rank lists of this rank's tiles, all in local memory
	const uint4* tileOrder;     // the work order: kTileClasses lists of nTiles {tile, list length, list offset, key-buffer flag} entries (tile_scan_kernel / shard_merge_heads_kernel)
	unsigned long long* keys;   // direct path: the frame's key buffer (see keyIndex) and its per-tile "holds fragments" map
	uint8_t* dirty;
};

// expanded per-triangle data kept in shared memory while a chunk of the bin list is processed
struct __align__(16) TriShared
{
	float x0, y0, x1, y1;
	float x2, y2, z0, z1;
	float z2, rcpArea, eAx, eAy;  // eA = v1 - v0; rcpArea = refined reciprocal of |area| (see divByArea)
	float eBx, eBy, eCx, eCy;     // eB = v2 - v1, eC = v0 - v2
	float area;                   // |Cross(v0-v1, v0-v2)|, CUDAROPs.cu:336
	uint32_t prim;
	int zLow;                     // conservative lower bound of any fragment's integer depth
	uint32_t yExtent;             // pixel-parallel path: bbox rows y0 | y1 << 16 for the warp-level reject
	float divLo;                  // weights >= divLo take the hoisted division; +inf disables it for this triangle
};

struct Bary { float b0, b1, b2, z; };

__device__ __forceinline__ bool topLeftEdge(float ex, float ey) { return ey > 0.f || (ey == 0.f && ex < 0.f); }

// IsPixelInTriangle + weights + depth of RasterizePixelAtomic (CUDAROPs.cu:128-145, 333-345, 386).
// Returns false if the fragment is not produced.  zi = cvt.rzi.s32.f32(z * 2^31).
template <bool TOPLEFT>
__device__ __forceinline__ bool fragmentTest(const TriShared& t, float fx, float fy, Bary& out, int& zi)
{
	const float w2 = diff2(fx - t.x0, t.eAy, fy - t.y0, t.eAx);
	const float w0 = diff2(fx - t.x1, t.eBy, fy - t.y1, t.eBx);
	const float w1 = diff2(fx - t.x2, t.eCy, fy - t.y2, t.eCx);
	if (!(w0 >= 0.f && w1 >= 0.f && w2 >= 0.f)) return false;
	const float wMin = fminf(fminf(w0, w1), w2);
	if (wMin >= t.divLo)
	{
		out.b0 = divByArea(w0, t.area, t.rcpArea); out.b1 = divByArea(w1, t.area, t.rcpArea); out.b2 = divByArea(w2, t.area, t.rcpArea);
	}
	else
	{
		// exact-edge fragments (some weight is 0) and degenerate operands: rare, fully general path
		if (TOPLEFT)
		{
			if ((w2 == 0.f && !topLeftEdge(t.eAx, t.eAy)) || (w0 == 0.f && !topLeftEdge(t.eBx, t.eBy)) ||
				(w1 == 0.f && !topLeftEdge(t.eCx, t.eCy)))
				return false;
		}
		out.b0 = __fdiv_rn(w0, t.area); out.b1 = __fdiv_rn(w1, t.area); out.b2 = __fdiv_rn(w2, t.area);
	}
	out.z = sum3(out.b0, t.z0, out.b1, t.z1, out.b2, t.z2);
	if (out.z < 0.f || out.z > 1.f) return false;
	zi = __float2int_rz(out.z * 2147483648.f);
	// a saturated INT_MAX never passes `zf < (float)INT_MAX` against the cleared buffer (CUDAROPs.cu:242)
	return zi != 0x7FFFFFFF;
}

__device__ __forceinline__ void expandRecord(const TriRecord& r, uint32_t prim, TriShared& t)
{
	t.x0 = r.x0; t.y0 = r.y0; t.x1 = r.x1; t.y1 = r.y1; t.x2 = r.x2; t.y2 = r.y2;
	t.z0 = r.z0; t.z1 = r.z1; t.z2 = r.z2; t.area = r.area;
	t.eAx = r.x1 - r.x0; t.eAy = r.y1 - r.y0;
	t.eBx = r.x2 - r.x1; t.eBy = r.y2 - r.y1;
	t.eCx = r.x0 - r.x2; t.eCy = r.y0 - r.y2;
	t.prim = prim;
	t.divLo = r.divLo; t.rcpArea = r.rcpArea; t.zLow = r.zLow;
	t.yExtent = 0;
}

// triangle setup: one thread per primitive of [primLo, primHi).
// cull (CUDARenderer.cu:760-777), frustum reject (CUDAROPs.cu:147-178), screen mapping (:88-97),
// inflated bbox (:180-200), bin rectangle (CUDARenderer.cu:935-942), per-tile counts.
//
// DIRECT: the direct path for small triangles.  The reference tests a triangle against every pixel of every bin of its
// rectangle (CUDARenderer.cu:1116-1127); the pixels that can pass are those of the triangle's candidate box
// (rasterConstants) inside the rectangle's bins.  When that is at most fp.directMaxArea pixels the thread runs the
// fragment test on them right here and resolves depth with a 64-bit atomicMin on the frame's key buffer - the same
// (z_int << 32 | primitive id) key and the same order-independent min the raster kernel keeps per tile - and the
// triangle is never counted, scattered, sorted or staged.  A triangle without any candidate pixel (between the pixel
// centres) costs nothing further.  The raster kernel merges the key buffer into its tiles' resolve (and resets it).
// On a scene of sub-pixel triangles this removes the bin lists altogether; the 66 MB key buffer of a 4K frame lives
// in B200's 126 MB L2.
//
// Structure.  A primitive costs two DEPENDENT memory round trips (vertex reference -> three positions) before a few
// hundred instructions of arithmetic, and a one-primitive-per-thread kernel idles on them (ncu: 28-45 % of the warp
// samples wait for the positions, issue slots ~50 % busy, whatever the occupancy).  So the kernel is persistent - every
// WARP strides over groups of 32 consecutive primitives - and every thread runs a three-stage pipeline of asynchronous
// copies (cp.async, global -> shared without passing through registers): while the warp works on group i, the positions
// of group i+1 and the vertex references of group i+2 are in flight.  A thread only ever reads back what it copied
// itself, so the pipeline needs no barrier.
// The direct path then runs in two phases with a stream compaction through shared memory in between, warp by warp (a
// CTA-wide compaction cost two barriers per 256 primitives and 32 % of the warp samples, because culled and visible
// warps differ widely in work).  Phase A is the setup proper, one lane per primitive: cull, rectangle, candidate box; a
// small triangle that has candidate pixels leaves a 48-byte record in the warp's shared-memory slice.  Phase B shares
// each record between k = 32 / records neighbouring lanes (a power of two) that take every k-th candidate pixel - the
// culled, off-pixel and binned primitives are gone, the lanes stay busy.
#ifndef B2R_SETUP_CTAS
#define B2R_SETUP_CTAS 5
#endif
struct __align__(16) DirectRecord
{
	float x0, y0, x1, y1;
	float x2, y2, z0, z1;
	float z2; uint32_t prim, boxX /*x0 | x1 << 16*/, boxY /*y0 | y1 << 16*/;
};
template <bool TOPLEFT, bool DIRECT>
__global__ void __launch_bounds__(kChunk, B2R_SETUP_CTAS) triangle_setup_kernel(
	const uint4* __restrict__ triRefs, const float4* __restrict__ pos4, uint2* __restrict__ rects,
	uint32_t* __restrict__ tileCount, uint32_t* __restrict__ bigList, uint32_t* __restrict__ counters /*[0]=visible,[1]=nBig,[2]=slots*/,
	uint32_t* __restrict__ binSlots /*DIRECT: the primitives bin_scatter_kernel has to visit, one 32-entry slot per warp that has any*/,
	unsigned long long* __restrict__ keys, uint8_t* __restrict__ dirty, unsigned long long* __restrict__ frameInfo,
	uint32_t primLo, uint32_t primHi, const __grid_constant__ FrameParams fp)
{
	pdlPrologue();
	__shared__ __align__(16) uint4 sRef[2][kChunk];              // vertex references of chunks i+1 / i+2
	__shared__ __align__(16) float4 sPos[2][3][kChunk];          // positions of chunks i / i+1
	__shared__ __align__(16) DirectRecord sRec[DIRECT ? kChunk : 1];      // one slot per thread, written by the thread itself
	__shared__ uint8_t sRecLane[DIRECT ? kChunk : 1];                      // per warp: the lanes that left a record, compacted
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	const uint32_t nChunks = (primHi - primLo + 31u) / 32u;        // ("chunk" = a warp's group of 32 primitives)
	const uint32_t firstChunk = blockIdx.x * (kChunk / 32) + warp;
	if (firstChunk >= nChunks) return;
	auto primOf = [&](uint32_t chunk) { return primLo + chunk * 32u + lane; };
	// pipeline fill (the step is gridDim.x chunks, and a chunk's buffers are chosen by the parity of its POSITION in this
	// CTA's sequence, not of its index)
	const uint32_t step = gridDim.x * (kChunk / 32);
	const SmemAddr aRef = smemAddress(&sRef[0][tid]), aPos = smemAddress(&sPos[0][0][tid]);      // this thread's slots: + buffer / vertex strides
	constexpr uint32_t kRefBuf = kChunk * (uint32_t)sizeof(uint4), kPosVtx = kChunk * (uint32_t)sizeof(float4), kPosBuf = 3u * kPosVtx;
	{
		const uint32_t c0 = firstChunk;
		if (primOf(c0) < primHi) cpAsync16(aRef, triRefs + primOf(c0));
		cpAsyncCommit(); cpAsyncWaitAll();
		if (primOf(c0) < primHi)
		{
			const uint4 tr = sRef[0][tid];
			cpAsync16(aPos, pos4 + tr.x); cpAsync16(aPos + kPosVtx, pos4 + tr.y); cpAsync16(aPos + 2u * kPosVtx, pos4 + tr.z);
		}
		if (c0 + step < nChunks && primOf(c0 + step) < primHi) cpAsync16(aRef + kRefBuf, triRefs + primOf(c0 + step));
		cpAsyncCommit();
	}
	uint32_t nFrags = 0, visAcc = 0, directAcc = 0;
	// statistics: a few atomics per warp, spread over kVisSlots padded slots (tile_scan_kernel sums them)
	auto flushStatistics = [&]()
	{
		uint32_t* slot = &counters[kVisBase + (firstChunk % kVisSlots) * 32];
		if (lane == 0 && visAcc) atomicAdd(slot, visAcc);
		if (DIRECT)
		{
			unsigned long long dt = directAcc;
			#pragma unroll
			for (int d = 16; d > 0; d >>= 1) dt += __shfl_down_sync(0xFFFFFFFFu, dt, d);
			if (lane == 0 && dt) atomicAdd(reinterpret_cast<unsigned long long*>(slot) + 1, dt);
		}
		visAcc = 0; directAcc = 0;
	};
	uint32_t it = 0;
	for (uint32_t chunk = firstChunk; chunk < nChunks; chunk += step, ++it)
	{
		const uint32_t cur = it & 1u, nxt = cur ^ 1u;
		uint32_t nRec = 0;
		// positions of this chunk and references of the next have landed (requested one whole iteration ago)
		cpAsyncWaitAll();
		// request: positions of the next chunk, references of the one after (into the buffers this iteration does not read)
		{
			const uint32_t c1 = chunk + step, c2 = chunk + 2u * step;
			float4 keep0, keep1, keep2;
			const uint32_t prim0 = primOf(chunk);
			// (this chunk's positions are read into registers first: sRef[cur] is about to be overwritten, sPos[cur] is not)
			keep0 = sPos[cur][0][tid]; keep1 = sPos[cur][1][tid]; keep2 = sPos[cur][2][tid];
			if (c1 < nChunks && primOf(c1) < primHi)
			{
				const uint4 tr = sRef[nxt][tid];
				const SmemAddr aNxt = aPos + nxt * kPosBuf;
				cpAsync16(aNxt, pos4 + tr.x); cpAsync16(aNxt + kPosVtx, pos4 + tr.y); cpAsync16(aNxt + 2u * kPosVtx, pos4 + tr.z);
			}
			if (c2 < nChunks && primOf(c2) < primHi) cpAsync16(aRef + cur * kRefBuf, triRefs + primOf(c2));
			cpAsyncCommit();

			// ---- phase A: setup of primitive prim0
			const float4 p0 = keep0, p1 = keep1, p2 = keep2;
			const uint32_t prim = prim0;
			bool visible = false;
			int bx0 = 1, bx1 = 0, by0 = 1, by1 = 0;
			uint32_t nTiles = 0, directTiles = 0;
			bool record = false;
			if (prim < primHi)
			{
				visible = true;
				if ((fp.cullMode & 0xFFu) != 2u)
				{
					const float ax = p1.x - p0.x, ay = p1.y - p0.y, az = p1.z - p0.z;
					const float bx = p2.x - p0.x, by = p2.y - p0.y, bz = p2.z - p0.z;
					const float nx = diff2(ay, bz, az, by), ny = diff2(az, bx, ax, bz), nz = diff2(ax, by, ay, bx);
					if (backFace(nx, ny, nz, fp.camFwd, fp.cullMode)) visible = false;
				}
				if (visible)
				{
					const bool xOut = (p0.x < -1.f && p1.x < -1.f && p2.x < -1.f) || (p0.x > 1.f && p1.x > 1.f && p2.x > 1.f);
					const bool yOut = (p0.y < -1.f && p1.y < -1.f && p2.y < -1.f) || (p0.y > 1.f && p1.y > 1.f && p2.y > 1.f);
					const bool zOut = (p0.z < 0.f && p1.z < 0.f && p2.z < 0.f) || (p0.z > 1.f && p1.z > 1.f && p2.z > 1.f);
					visible = !(xOut || yOut || zOut);
				}
				if (visible)
				{
					TriRecord r;
					const ScreenBounds b = screenMap(p0, p1, p2, fp, r);
					const int xMin = bboxCoord(b.minx, -1, (int)fp.width);
					const int yMin = bboxCoord(b.miny, -1, (int)fp.height);
					const int xMax = bboxCoord(b.maxx, +1, (int)fp.width);
					const int yMax = bboxCoord(b.maxy, +1, (int)fp.height);
					// C integer division of the non-negative shorts (CUDARenderer.cu:935-942); a shift for power-of-two bins
					const bool px = fp.binShiftX >= 0, py = fp.binShiftY >= 0;
					bx0 = clampi(px ? xMin >> fp.binShiftX : xMin / (int)fp.binW, 0, (int)fp.binsX);
					by0 = clampi(py ? yMin >> fp.binShiftY : yMin / (int)fp.binH, 0, (int)fp.binsY);
					bx1 = clampi(px ? xMax >> fp.binShiftX : xMax / (int)fp.binW, 0, (int)fp.binsX - 1);
					by1 = clampi(py ? yMax >> fp.binShiftY : yMax / (int)fp.binH, 0, (int)fp.binsY - 1);
					if (bx0 <= bx1 && by0 <= by1)
					{
						nTiles = (uint32_t)(bx1 - bx0 + 1) * (uint32_t)(by1 - by0 + 1);
						if (DIRECT)
						{
							rasterConstants(b, fp, r);
							// candidate pixels: the candidate box inside the bins of the rectangle (the box is already clamped to the screen)
							const int x0 = max((int)(r.cbX & 0xFFFFu), bx0 * (int)fp.binW), x1 = min((int)(r.cbX >> 16), (bx1 + 1) * (int)fp.binW);
							const int y0 = max((int)(r.cbY & 0xFFFFu), by0 * (int)fp.binH), y1 = min((int)(r.cbY >> 16), (by1 + 1) * (int)fp.binH);
							const uint32_t area = (x1 > x0 && y1 > y0) ? (uint32_t)(x1 - x0) * (uint32_t)(y1 - y0) : 0u;
							if (area <= fp.directMaxArea)
							{
								directTiles = nTiles;
								nTiles = 0;
								if (area)
								{
									// the record goes straight into the thread's own slot (three 16-byte stores: a DirectRecord
									// assembled in a local variable and copied after the vote lived in local memory)
									record = true;
									float4* q = reinterpret_cast<float4*>(&sRec[tid]);
									q[0] = make_float4(r.x0, r.y0, r.x1, r.y1);
									q[1] = make_float4(r.x2, r.y2, r.z0, r.z1);
									q[2] = make_float4(r.z2, __uint_as_float(prim), __uint_as_float((uint32_t)x0 | ((uint32_t)x1 << 16)), __uint_as_float((uint32_t)y0 | ((uint32_t)y1 << 16)));
								}
							}
						}
					}
				}
				// culled primitives are marked (0xFFFF, 0): an empty rectangle that b2r_read_visible can tell
				// from a visible primitive whose rectangle happens to be empty
				// (bit 31 of the first word: resolved here, not to be scattered into the lists)
				rects[prim] = visible ? make_uint2((uint32_t)bx0 | ((uint32_t)bx1 << 16) | (directTiles ? kRectDirect : 0u), (uint32_t)by0 | ((uint32_t)by1 << 16))
				                      : make_uint2(0x0000FFFFu, 0x0000FFFFu);
			}
			if (DIRECT)
			{
				// compaction of the lanes that left a record
				const unsigned m = __ballot_sync(0xFFFFFFFFu, record);
				nRec = (uint32_t)__popc(m);
				if (record) sRecLane[warp * 32u + (uint32_t)__popc(m & ((1u << lane) - 1u))] = (uint8_t)lane;
			}
			// statistics of the group, accumulated per warp and flushed when the warp is done (and every 64 groups, so that
			// a lane's 32-bit sum of rectangles of at most 2^22 tiles cannot wrap)
			visAcc += (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, visible));
			if (DIRECT) directAcc += min(directTiles, 0x00FFFFFFu);
			if ((it & 63u) == 63u) flushStatistics();
			// per-tile counts, warp-aggregated
			if (nTiles > fp.bigRectTiles)
			{
				bigList[atomicAdd(&counters[1], 1u)] = prim;      // handled cooperatively (binBigBody)
				nTiles = 0;
			}
			if (DIRECT)
			{
				// the few triangles that still go into the lists: the scatter kernel visits them through 32-entry slots (this warp's
				// ids in lane order, kNoPrim where a lane has none) instead of re-reading the rectangle of every primitive
				const unsigned binned = __ballot_sync(0xFFFFFFFFu, nTiles > 0u);
				if (binned)
				{
					uint32_t slot = 0;
					if (lane == 0) slot = atomicAdd(&counters[kBinSlotsWord], 1u);
					slot = __shfl_sync(0xFFFFFFFFu, slot, 0);
					binSlots[(size_t)slot * 32u + lane] = nTiles > 0u ? prim : kNoPrim;
				}
			}
			warpBinRects<false>(nTiles, bx0, bx1, by0, by1, prim, tileCount, nullptr, nullptr, fp);
		}
		if (!DIRECT) continue;

		// ---- phase B: the recorded triangles of the group
		__syncwarp();
		if (nRec)
		{
			// lanes per record: the largest power of two <= 32 / records, 2^(5 - ceil(log2(records)))
			const uint32_t shareLog = (uint32_t)__clz(nRec - 1u) - 27u;
			const uint32_t share = 1u << shareLog;
			const uint32_t recIdx = lane >> shareLog, first = lane & (share - 1u);
			if (recIdx < nRec)
			{
				const DirectRecord& q = sRec[warp * 32u + sRecLane[warp * 32u + recIdx]];
				const int x0 = (int)(q.boxX & 0xFFFFu), x1 = (int)(q.boxX >> 16), y0 = (int)(q.boxY & 0xFFFFu), y1 = (int)(q.boxY >> 16);
				const uint32_t w = (uint32_t)(x1 - x0), area = w * (uint32_t)(y1 - y0);
				if (first < area)
				{
					TriShared t;
					t.x0 = q.x0; t.y0 = q.y0; t.x1 = q.x1; t.y1 = q.y1; t.x2 = q.x2; t.y2 = q.y2;
					t.z0 = q.z0; t.z1 = q.z1; t.z2 = q.z2;
					// the same functions of the same coordinates as rasterConstants / expandRecord: same bits as in the raster kernel
					t.area = fabsf(diff2(t.x0 - t.x1, t.y0 - t.y2, t.y0 - t.y1, t.x0 - t.x2));
					t.rcpArea = refinedReciprocal(t.area);
					t.divLo = divisionThreshold(t.area);
					t.eAx = q.x1 - q.x0; t.eAy = q.y1 - q.y0;
					t.eBx = q.x2 - q.x1; t.eBy = q.y2 - q.y1;
					t.eCx = q.x0 - q.x2; t.eCy = q.y0 - q.y2;
					t.prim = q.prim; t.zLow = 0; t.yExtent = 0;
					const bool px = fp.binShiftX >= 0, py = fp.binShiftY >= 0;
					const bool replicated = fp.ownerCount > 1u && !fp.shardGeometry;      // sort-first with a replicated front end: own tiles only
					const bool tileIsBlock = fp.binShiftX == 4 && fp.binShiftY == 4 && fp.keyBlocksX == fp.binsX;
					int y = y0, x = x0 + (int)first;
					while (x >= x1) { x -= (int)w; ++y; }
					for (uint32_t i = first; i < area; i += share)
					{
						Bary bc; int zi;
						if (fragmentTest<TOPLEFT>(t, (float)x, (float)y, bc, zi))
						{
							const size_t key = keyIndex(fp.keyBlocksX, x, y);
							// (16x16 bins on the key buffer's block grid: the tile is the key's block)
							const uint32_t tile = tileIsBlock ? (uint32_t)(key >> 8)
								: (uint32_t)(px ? x >> fp.binShiftX : x / (int)fp.binW) + (uint32_t)(py ? y >> fp.binShiftY : y / (int)fp.binH) * fp.binsX;
							if (!replicated || tile % fp.ownerCount == fp.ownerRank)
							{
								atomicMin(keys + key, ((unsigned long long)(uint32_t)zi << 32) | t.prim);
								dirty[tile] = 1;
								++nFrags;
							}
						}
						x += (int)share;
						while (x >= x1) { x -= (int)w; ++y; }
					}
				}
			}
		}
		__syncwarp();                      // every lane is done with the records before the next group's are written
	}
	flushStatistics();
	if (DIRECT && fp.countFragments)
	{
		nFrags = __reduce_add_sync(0xFFFFFFFFu, nFrags);
		if (lane == 0 && nFrags) atomicAdd(&frameInfo[4], (unsigned long long)nFrags);
	}
}

// ---- texture sampling (CUDATextureSampler.cuh:80-191) -------------------------------------
struct RGB { float r, g, b; };

__device__ __forceinline__ bool texelCoords(const TexDesc& t, float u, float v, int& x, int& y)
{
	x = __float2int_rz(fmaf(u, (float)t.w, 0.5f));
	y = __float2int_rz(fmaf(v, (float)t.h, 0.5f));
	return !(x < 0 || y < 0 || (uint32_t)x >= t.w || (uint32_t)y >= t.h);
}
__device__ __forceinline__ RGB samplePoint(const TexDesc& t, float u, float v)
{
	int x, y;
	if (!texelCoords(t, u, v, x, y)) return RGB{ 1.f, 0.f, 1.f };
	const uint32_t s = __ldg(t.px + texelIndex(t, (uint32_t)x + (uint32_t)y * t.w));
	return RGB{ __fdiv_rn((float)(s & 0xFFu), 255.f), __fdiv_rn((float)((s >> 8) & 0xFFu), 255.f), __fdiv_rn((float)((s >> 16) & 0xFFu), 255.f) };
}
__device__ __forceinline__ RGB sampleLinear(const TexDesc& t, float u, float v)
{
	int x, y;
	if (!texelCoords(t, u, v, x, y)) return RGB{ 1.f, 0.f, 1.f };
	const int texSize = (int)(t.w * t.h);
	const int o = x + y * (int)t.w;
	int n0 = o - 1;          if (n0 < 0) n0 = o;
	int n1 = o + 1;          if (n1 >= texSize) n1 = o;
	int n2 = o + (int)t.w;   if (n2 >= texSize) n2 = o;
	int n3 = o - (int)t.w;   if (n3 < 0) n3 = o;
	const uint32_t s0 = __ldg(t.px + texelIndex(t, (uint32_t)n0)), s1 = __ldg(t.px + texelIndex(t, (uint32_t)n1)), s2 = __ldg(t.px + texelIndex(t, (uint32_t)n2)),
		s3 = __ldg(t.px + texelIndex(t, (uint32_t)n3)), sc = __ldg(t.px + texelIndex(t, (uint32_t)o));
	// all partial sums are exact in fp32 (multiples of 1/8 below 2^11), so the order is free
	const float r = (float)((s0 & 0xFFu) + (s1 & 0xFFu) + (s2 & 0xFFu) + (s3 & 0xFFu)) * 0.125f + (float)(sc & 0xFFu) * 0.5f;
	const float g = (float)(((s0 >> 8) & 0xFFu) + ((s1 >> 8) & 0xFFu) + ((s2 >> 8) & 0xFFu) + ((s3 >> 8) & 0xFFu)) * 0.125f + (float)((sc >> 8) & 0xFFu) * 0.5f;
	const float b = (float)(((s0 >> 16) & 0xFFu) + ((s1 >> 16) & 0xFFu) + ((s2 >> 16) & 0xFFu) + ((s3 >> 16) & 0xFFu)) * 0.125f + (float)((sc >> 16) & 0xFFu) * 0.5f;
	const float rev = __frcp_rn(255.f);
	return RGB{ r * rev, g * rev, b * rev };
}
__device__ __forceinline__ RGB sampleTex(const TexDesc& t, float u, float v, uint32_t sampleState)
{
	return sampleState == 1u ? sampleLinear(t, u, v) : samplePoint(t, u, v);
}

// RGBRaw.h:23-32: truncating pack to 0x00RRGGBB
__device__ __forceinline__ uint32_t packColour(float r, float g, float b)
{
	const uint32_t b0 = __float2uint_rz(b * 255.f) & 0xFFu;
	const uint32_t b1 = __float2uint_rz(g * 255.f) & 0xFFu;
	const uint32_t b2 = __float2uint_rz(r * 255.f) & 0xFFu;
	return b0 | (b1 << 8) | (b2 << 16);
}

// attribute interpolation S(a) = w0*(a0*iw0) + w1*(a1*iw1) + w2*(a2*iw2), CUDAROPs.cu:50-84
struct Interp
{
	float b0, b1, b2, iw0, iw1, iw2;
	__device__ __forceinline__ float operator()(float a0, float a1, float a2) const
	{
		return sum3(b0, a0 * iw0, b1, a1 * iw1, b2, a2 * iw2);
	}
};

// ShadePixelSafe, CUDAROPs.cu:411-470, with the contraction pattern of the reference build
__device__ __forceinline__ RGB shadePixel(const RasterArgs& A, const MeshConst& mc, uint32_t sampleState,
	float u, float v, float nx, float ny, float nz, float tx, float ty, float tz, float vx, float vy, float vz)
{
	const RGB D = sampleTex(A.textures[mc.tex[0]], u, v, sampleState);
	if (mc.tex[1] < 0) return D;
	const RGB Ns = sampleTex(A.textures[mc.tex[1]], u, v, sampleState);
	const float cx = diff2(ty, nz, tz, ny), cy = diff2(tz, nx, tx, nz), cz = diff2(tx, ny, ty, nx);   // Cross(tan, n); binormal = -C
	const float fx = fmaf(Ns.r, 2.f, -1.f), fy = fmaf(Ns.g, 2.f, -1.f), fz = fmaf(Ns.b, 2.f, -1.f);
	const float Nx = fmaf(fz, nx, fmaf(tx, fx, -(cx * fy)));
	const float Ny = fmaf(fz, ny, fmaf(ty, fx, -(cy * fy)));
	const float Nz = fmaf(fz, nz, fmaf(tz, fx, -(cz * fy)));
	const float L = 0.577f;
	const float nxL = Nx * L;
	float oa = fmaf(Nz, L, fmaf(Ny, L, -nxL));       // Dot(-N, lightDirection); ClampFast(oa,0,oa) is the identity
	oa = __fdiv_rn(oa, 3.14159274101257324f);
	oa = oa * 7.0f;
	RGB c{ D.r * oa, D.g * oa, D.b * oa };
	if (mc.tex[2] < 0 || mc.tex[3] < 0) return c;
	const RGB S = sampleTex(A.textures[mc.tex[2]], u, v, sampleState);
	const RGB G = sampleTex(A.textures[mc.tex[3]], u, v, sampleState);
	float dotLN = fmaf(Nz, -L, fmaf(Ny, -L, nxL));
	dotLN = dotLN + dotLN;
	const float Rx = fmaf(Nx, dotLN, -L), Ry = fmaf(Ny, dotLN, L), Rz = fmaf(Nz, dotLN, L);
	float angle = fmaf(Rz, vz, fmaf(Rx, vx, vy * Ry));
	angle = (angle < 0.f) ? 0.f : angle;
	angle = (angle > 1.f) ? 1.f : angle;
	angle = powf(angle, G.r * 25.f);
	c.r = fmaf(angle, S.r, c.r + 0.05f);
	c.g = fmaf(angle, S.g, c.g + 0.05f);
	c.b = fmaf(angle, S.b, c.b + 0.05f);
	const float gb = (c.g < c.b) ? c.b : c.g;
	const float mx = (c.r < gb) ? gb : c.r;
	if (mx > 1.f)
	{
		const float rev = __frcp_rn(mx);
		c.r *= rev; c.g *= rev; c.b *= rev;
	}
	return c;
}

constexpr int kListChunk = 256;       // bin-list entries processed per sweep iteration (one per thread)
#ifndef B2R_RASTER_LOOKAHEAD
#define B2R_RASTER_LOOKAHEAD 1        // a thread holds the vertex reference of its NEXT list entry in registers while it works on the current one
#endif
#ifndef B2R_SMALL_AREA
#define B2R_SMALL_AREA 48
#endif
constexpr int kSmallArea = B2R_SMALL_AREA;   // bbox-in-tile areas up to this are rasterised by one thread

__device__ __forceinline__ void prefetchL2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

// TEXTURED = false is the light variant for frames in which no mesh has a diffuse map and the
// visualisation is PBR (vertex colours only): the texture/normal-mapping path is compiled out, which
// frees enough registers for 6 resident CTAs per SM instead of 4.
#ifndef B2R_RASTER_LIGHT_CTAS
#define B2R_RASTER_LIGHT_CTAS 6
#endif
#ifndef B2R_RASTER_FULL_CTAS
#define B2R_RASTER_FULL_CTAS 4
#endif
// (the light instantiation that transforms vertices itself - sliced vertex stage of the sharded front end - spills
// ~180 bytes at the light register budget: B2R_RASTER_OTF_CTAS lets a build trade occupancy for registers)
#ifndef B2R_RASTER_OTF_CTAS
#define B2R_RASTER_OTF_CTAS 4      // measured at 4 ranks (gpurun_out/s3_multi4): 7 429 frames/s at 4 CTAs (59 registers, no spills), 7 384 at 5, 7 151 at 6
#endif
constexpr int rasterCtasPerSM(bool textured, bool otf) { return textured ? B2R_RASTER_FULL_CTAS : otf ? B2R_RASTER_OTF_CTAS : B2R_RASTER_LIGHT_CTAS; }

// big triangles of the pixel-parallel path as they wait in shared memory (64 bytes; the edge vectors are rebuilt from
// the vertices with the subtractions of expandRecord, hence the same bits)
struct __align__(16) TriBig
{
	float x0, y0, x1, y1;
	float x2, y2, z0, z1;
	float z2, area, rcpArea, divLo;
	uint32_t prim; int zLow; uint32_t yExtent, pad;
};
__device__ __forceinline__ void packBig(const TriShared& t, uint32_t yExtent, TriBig& b)
{
	b.x0 = t.x0; b.y0 = t.y0; b.x1 = t.x1; b.y1 = t.y1; b.x2 = t.x2; b.y2 = t.y2; b.z0 = t.z0; b.z1 = t.z1;
	b.z2 = t.z2; b.area = t.area; b.rcpArea = t.rcpArea; b.divLo = t.divLo; b.prim = t.prim; b.zLow = t.zLow; b.yExtent = yExtent; b.pad = 0;
}
__device__ __forceinline__ void unpackBig(const TriBig& b, TriShared& t)
{
	t.x0 = b.x0; t.y0 = b.y0; t.x1 = b.x1; t.y1 = b.y1; t.x2 = b.x2; t.y2 = b.y2; t.z0 = b.z0; t.z1 = b.z1;
	t.z2 = b.z2; t.area = b.area; t.rcpArea = b.rcpArea; t.divLo = b.divLo; t.prim = b.prim; t.zLow = b.zLow; t.yExtent = b.yExtent;
	t.eAx = b.x1 - b.x0; t.eAy = b.y1 - b.y0;
	t.eBx = b.x2 - b.x1; t.eBy = b.y2 - b.y1;
	t.eCx = b.x0 - b.x2; t.eCy = b.y0 - b.y2;
}

// the list of the tile a CTA rasterises
struct WorkItem
{
	uint32_t count, offset;           // length of the tile's list, its offset in binIds (single GPU / replicated front end)
	uint32_t segEnd[kMaxRanks], segOff[kMaxRanks];    // sharded front end: the ranks' segments of the list (prefix ends, offsets)
};

// global slot of entry e of a tile's (concatenated) list
__device__ __forceinline__ uint32_t* listSlotOf(const RasterArgs& A, const WorkItem& item, bool shard, uint32_t e)
{
	if (shard)
	{
		uint32_t seg = 0;
		while (e >= item.segEnd[seg]) ++seg;
		return const_cast<uint32_t*>(A.shard.binIds[seg]) + item.segOff[seg] + (e - (seg ? item.segEnd[seg - 1] : 0u));
	}
	return const_cast<uint32_t*>(A.binIds) + item.offset + e;
}

// The single-warp / whole-CTA sort helpers are real functions (not inlined): they run once per tile, and their
// registers must not count against the sweep's.
// warp 0: canonical order of the staged list (stored to global memory); returns warpSortTileList's verdict
__device__ B2R_NOINLINE int sortStagedList(const RasterArgs& A, const WorkItem& item, bool shard, const uint32_t* K, uint32_t n, WarpSortScratch& S)
{
	return warpSortTileList(K, n, S, threadIdx.x & 31u, [&](uint32_t pos, uint32_t key) { *listSlotOf(A, item, shard, pos) = key; });
}
// all threads: the bitonic fallback for lists whose runs warp 0 could not order
__device__ B2R_NOINLINE void sortStagedListFallback(const RasterArgs& A, const WorkItem& item, bool shard, uint32_t* K, uint32_t n)
{
	ctaBitonicSortList(K, n, [&](uint32_t pos, uint32_t key) { *listSlotOf(A, item, shard, pos) = key; });
}

// One CTA = one work item = one 16x16 pixel block (a tile, or one sub-block of a bin larger than 16x16), or kEmptyBatch
// blocks of empty tiles.  The items are numbered in the order tile_scan_kernel (sharded front end:
// shard_merge_heads_kernel) prepared - longest lists first, the empty tiles last - and the hardware hands out CTAs in
// index order, so the long lists start first and the tail of the kernel is made of short ones.  The grid covers the
// worst case (one item per block); the CTAs behind the last item leave at once.
// Lists of up to kStage entries are staged in shared memory; warp 0 then puts the staged list into canonical order
// (and stores it) while warps 1..7 already sweep it - the sweep does not depend on the order.
// the three transformed positions of a primitive.  OTF (sharded front end with a sliced vertex stage: a rank has transformed
// only the vertices of its own primitive slice): from the resident model-space positions with the vertex stage's own
// function - same bits as the stream vertex_transform_kernel writes.
template <bool OTF>
__device__ __forceinline__ void fetchPositions(const RasterArgs& A, const uint4& tr, float4& p0, float4& p1, float4& p2)
{
	if (OTF)
	{
		const MeshConst& mc = A.meshes[tr.w];
		const uint32_t d = mc.vtxInBase - mc.vtxOutBase;     // transformed-stream index -> IVertex record (mod 2^32)
		const float4 a = __ldg(A.posIn + (tr.x + d)), b = __ldg(A.posIn + (tr.y + d)), c = __ldg(A.posIn + (tr.z + d));
		float m[16];
		#pragma unroll
		for (int i = 0; i < 16; ++i) m[i] = __ldg(&mc.wvp[i]);
		p0 = transformPosition(m, a.x, a.y, a.z); p1 = transformPosition(m, b.x, b.y, b.z); p2 = transformPosition(m, c.x, c.y, c.z);
	}
	else
	{
		p0 = __ldg(A.pos4 + tr.x); p1 = __ldg(A.pos4 + tr.y); p2 = __ldg(A.pos4 + tr.z);
	}
}

// One work item of the raster kernel, by the whole CTA.  False: the frame has no item of that index (or is void).
template <bool TOPLEFT, bool COUNT, bool TEXTURED, bool OTF>
__device__ __forceinline__ bool rasterWorkItem(const RasterArgs& A, const FrameParams& fp, const uint32_t itemIdx)
{
	__shared__ __align__(16) TriBig sBig[kListChunk];   // the pixel-parallel triangles of the block
	__shared__ __align__(16) uint32_t sIds[kStage];    // the tile's list (primitive ids), staged for the sort and the sweep
	__shared__ unsigned long long sKey[kTile * kTile]; // depth/winner of the triangle-parallel path
	__shared__ WarpSortScratch sSort;
	__shared__ WorkItem sItem;
	__shared__ uint32_t sNumBig[2];                    // big-triangle count of the current / the next list chunk
	__shared__ uint32_t sSortFallback;

	if (A.frameInfo[1]) return false;      // list-capacity overflow: frame is re-run by the host
	const int tid = threadIdx.x;
	const uint32_t lane = (uint32_t)tid & 31u, warp = (uint32_t)tid >> 5;
	const uint32_t nTilesAll = fp.binsX * fp.binsY;
	const uint32_t subs = fp.subX * fp.subY;
	// this CTA's work item: class, first block, number of blocks (the class counts are the same few words for every
	// CTA: L1 hits but for the first CTA of an SM)
	uint32_t cls = 0, start = 0, run = 0, emptyBlocks = 0;
	#pragma unroll
	for (uint32_t c = 0; c < kTileClasses; ++c)
	{
		const uint32_t blocks = (uint32_t)__ldg(A.frameInfo + kInfoClassBase + c) * subs;
		if (itemIdx >= run) { cls = c; start = run; }
		if (c + 1 == kTileClasses) emptyBlocks = blocks;
		run += (c + 1 < kTileClasses) ? blocks : (blocks + kEmptyBatch - 1u) / kEmptyBatch;
	}
	if (itemIdx >= run) return false;
	uint32_t v0 = itemIdx - start, v1 = v0 + 1;
	if (cls == kTileClasses - 1u)
	{
		v0 *= kEmptyBatch;
		v1 = min(v0 + kEmptyBatch, emptyBlocks);
	}
	const bool shard = fp.shardGeometry != 0u;
	const WorkItem& item = sItem;

	// pixel block `sub` of bin `bin`: origin and extent inside the bin and the screen
	auto blockRect = [&](uint32_t bin, uint32_t sub, int& ox, int& oy, int& ex, int& ey)
	{
		uint32_t binX, subX;
		const uint32_t binY = divMagic(bin, fp.binsX, fp.binsXMagic, binX);
		const uint32_t subY = divMagic(sub, fp.subX, fp.subXMagic, subX);
		ox = (int)(binX * fp.binW + subX * kTile); oy = (int)(binY * fp.binH + subY * kTile);
		ex = min(min(ox + kTile, (int)((binX + 1) * fp.binW)), (int)fp.width);
		ey = min(min(oy + kTile, (int)((binY + 1) * fp.binH)), (int)fp.height);
	};
	// ROP write-back: row-major surface, or (sort-first) the rank's packed tile-major buffer that is
	// gathered with one contiguous NCCL transfer / written straight into a peer's memory
	auto storePixel = [&](uint32_t bin, uint32_t sub, uint32_t inBlock /*row * 16 + column*/, size_t pix, uint32_t colour, int depthOut, float zOut, uint32_t prim)
	{
		if (fp.packedOutput) A.framebuffer[((size_t)((bin - fp.ownerRank) / fp.ownerCount) * subs + sub) * (kTile * kTile) + inBlock] = colour;
		else A.framebuffer[pix] = colour;
		A.depth[pix] = depthOut;
		if (fp.debugSurfaces)
		{
			A.zSurface[pix] = zOut;
			A.winnerSurface[pix] = prim;
		}
	};
	if (cls == kTileClasses - 1u)
	{
		// a batch of empty tiles: untouched fragments = zero-cleared PixelShade whose colour32 is the clear colour
		// (Depth: (0-.985)/.015*255 < 0 -> 0 ; Normal: n = 0).  A warp clears whole blocks, kEmptyBatch / 8 of them, and
		// reads all their descriptors before the first store: the batch costs one memory round trip, not one per block
		// (with one block after the other per CTA, a 4K frame's empty tiles took a third of this kernel's CTA-slot time)
		const uint32_t clear = (fp.visState == 0u) ? packColour(fp.clearRGB[0], fp.clearRGB[1], fp.clearRGB[2]) : 0u;
		constexpr uint32_t kPerWarp = kEmptyBatch / (kTile * kTile / 32);
		uint32_t bins[kPerWarp], subIdx[kPerWarp];
		#pragma unroll
		for (uint32_t k = 0; k < kPerWarp; ++k)
		{
			const uint32_t vb = v0 + warp + k * (kTile * kTile / 32);
			bins[k] = kNoPrim; subIdx[k] = 0;
			if (vb < v1) bins[k] = __ldg(&A.tileOrder[(size_t)cls * nTilesAll + divMagic(vb, subs, fp.subsMagic, subIdx[k])].x);
		}
		#pragma unroll
		for (uint32_t k = 0; k < kPerWarp; ++k)
		{
			if (bins[k] == kNoPrim) continue;      // (warp-uniform)
			int ox, oy, ex, ey;
			blockRect(bins[k], subIdx[k], ox, oy, ex, ey);
			const int px = ox + (int)(lane & (kTile - 1));
			#pragma unroll
			for (uint32_t j = 0; j < kTile / 2; ++j)
			{
				const uint32_t row = (lane >> 4) + 2u * j;
				const int py = oy + (int)row;
				if (px < ex && py < ey) storePixel(bins[k], subIdx[k], row * kTile + (lane & (kTile - 1)), (size_t)px + (size_t)py * fp.width, clear, 0x7FFFFFFF, 0.f, kNoPrim);
			}
		}
		return true;
	}
	{
		uint32_t sub;
		const uint4 desc = __ldg(A.tileOrder + (size_t)cls * nTilesAll + divMagic(v0, subs, fp.subsMagic, sub));      // {tile, list length, list offset, 0}
		const uint32_t bin = desc.x;
		int ox, oy, ex, ey;
		blockRect(bin, sub, ox, oy, ex, ey);
		const int px = ox + (tid & (kTile - 1)), py = oy + (tid >> 4);
		const bool live = px < ex && py < ey;
		const float fx = (float)px, fy = (float)py;
		// rows covered by this warp (two 16-pixel rows): triangles whose bbox misses them are skipped warp-wide
		const int warpY0 = oy + ((tid >> 5) << 1), warpY1 = warpY0 + 2;

		int bestZ = 0x7FFFFFFF;          // register copy of this pixel's depth/winner for the pixel-parallel path
		uint32_t bestPrim = kNoPrim;
		uint32_t nFrags = 0;             // fragments that pass the inside + z-range tests (COUNT builds only)
		const uint32_t n = desc.y;

		auto listSlot = [&](uint32_t e) -> uint32_t* { return listSlotOf(A, item, shard, e); };
		// One list entry -> screen-space triangle + its candidate box inside this CTA's block.  Returns the number of
		// candidate pixels (0: nothing to do).
		auto expandEntry = [&](uint32_t prim, TriShared& t, int& x0, int& x1, int& y0, int& y1) -> int
		{
			// rebuild the screen-space triangle from the static reference + the transformed positions
			// (same code, hence the same bits, wherever a triangle is expanded)
			const uint4 tr = __ldg(A.triRefs + prim);
			float4 p0, p1, p2;
			fetchPositions<OTF>(A, tr, p0, p1, p2);
			TriRecord r;
			rasterConstants(screenMap(p0, p1, p2, fp, r), fp, r);
			// candidate pixels = the record's conservative box clipped to this CTA's block; the box
			// contains every pixel whose edge functions can all be >= 0, so this equals the
			// reference's test of every pixel of the bin (CUDARenderer.cu:1116-1127)
			x0 = max((int)(r.cbX & 0xFFFFu), ox); x1 = min((int)(r.cbX >> 16), ex);
			y0 = max((int)(r.cbY & 0xFFFFu), oy); y1 = min((int)(r.cbY >> 16), ey);
			const int area = (x1 > x0 && y1 > y0) ? (x1 - x0) * (y1 - y0) : 0;
			if (area > 0) expandRecord(r, prim, t);
			return area;
		};
		// pixel-parallel path: every thread tests its own pixel against the large triangles staged in sBig
		auto pixelParallel = [&](uint32_t nBig)
		{
			for (uint32_t j = 0; j < nBig; ++j)
			{
				const uint32_t yExtent = sBig[j].yExtent;
				const int ty0 = (int)(yExtent & 0xFFFFu), ty1 = (int)(yExtent >> 16);
				if (ty1 <= warpY0 || ty0 >= warpY1) continue;          // warp-uniform
				if (!COUNT && sBig[j].zLow > bestZ) continue;          // cannot win this pixel (per lane)
				TriShared t;
				unpackBig(sBig[j], t);
				Bary bc; int zi;
				if (live && fragmentTest<TOPLEFT>(t, fx, fy, bc, zi))
				{
					if (COUNT) ++nFrags;
					if (zi < bestZ || (zi == bestZ && t.prim < bestPrim)) { bestZ = zi; bestPrim = t.prim; }
				}
			}
		};

		unsigned long long best = kEmptyKey;
		if (n)
		{
			sKey[tid] = kEmptyKey;
			if (tid == 0) { sNumBig[0] = 0; sNumBig[1] = 0; sSortFallback = 0; sItem.count = n; sItem.offset = desc.z; }
			if (shard)
			{
				// Sharded front end.  Rank p built the list of ITS primitive slice for every tile and pushed the heads
				// (after its scan) and the ids (after its scatter) of the tiles owned here into this rank's memory.  The
				// slices are contiguous and ordered by rank, so the tile's list is the concatenation of the segments in
				// rank order: one lane per rank reads that rank's head.
				if (tid < 32)
				{
					uint32_t cnt = 0, off = 0;
					if (lane < fp.ownerCount)
					{
						cnt = __ldg(A.shard.tileCount[lane] + bin);
						off = __ldg(A.shard.tileOffset[lane] + bin);
					}
					uint32_t incl = cnt;
					#pragma unroll
					for (int d = 1; d < kMaxRanks; d <<= 1)
					{
						const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
						if (lane >= (uint32_t)d) incl += t;
					}
					if (lane < kMaxRanks) { sItem.segEnd[lane] = incl; sItem.segOff[lane] = off; }
				}
				__syncthreads();
			}
			// lists of up to kStage entries are staged in shared memory: warp 0 sorts them there (canonical ascending order
			// in global memory - the sweep itself does not depend on the order) while the other warps sweep them
			const bool staged = n <= (uint32_t)kStage;
			if (staged)
			{
				const uint32_t* list = A.binIds + desc.z;
				for (uint32_t e = tid; e < n; e += kListChunk)
				{
					const uint32_t prim = shard ? *listSlot(e) : __ldg(list + e);
					sIds[e] = prim;
					if (e < (uint32_t)kListChunk) prefetchL2(A.triRefs + prim);      // the vertex references of the first entries
				}
			}
			__syncthreads();                  // sKey / counters initialised, list staged
			const bool sortHere = staged && fp.sortLists && n > 1u;
			if (warp == 0 && sortHere)
			{
				const int how = sortStagedList(A, item, shard, sIds, n, sSort);
				if (how == 2 && lane == 0) sSortFallback = 1u;
			}
			auto idAt = [&](uint32_t e) -> uint32_t { return staged ? sIds[e] : *listSlot(e); };

			// Sweep 1, no barriers: every sweeping thread walks its entries of the list.  Small triangles are resolved
			// right away (triangle-parallel, atomicMin into sKey); large ones are only collected in sBig - up to kListChunk of
			// them for the WHOLE list - and rasterised pixel-parallel after the sweep.  The warps of the CTA therefore drift
			// apart over a long list instead of meeting at two barriers per chunk, and their memory round trips
			// (vertex reference -> positions) overlap each other's arithmetic.  While warp 0 sorts, warps 1..7 sweep.
			const uint32_t sweepers = sortHere ? (uint32_t)kListChunk - 32u : (uint32_t)kListChunk;
			const uint32_t me = sortHere ? (uint32_t)tid - 32u : (uint32_t)tid;      // warp 0: >= sweepers while it sorts
			if (me < sweepers)
				for (uint32_t e = me; e < n; e += sweepers)
				{
					const uint32_t prim = idAt(e);
					if (staged && e + sweepers < n) prefetchL2(A.triRefs + sIds[e + sweepers]);
					TriShared t;
					int x0, x1, y0, y1;
					const int area = expandEntry(prim, t, x0, x1, y0, y1);
					if (area > kSmallArea)
					{
						const uint32_t slotBig = atomicAdd(&sNumBig[0], 1u);
						if (slotBig < (uint32_t)kListChunk) packBig(t, (uint32_t)y0 | ((uint32_t)y1 << 16), sBig[slotBig]);   // bbox rows for the warp-level reject
					}
					else if (area > 0)
					{
						// triangle-parallel: this thread walks the few candidate pixels
						for (int y = y0; y < y1; ++y)
							for (int x = x0; x < x1; ++x)
							{
								Bary bc; int zi;
								if (fragmentTest<TOPLEFT>(t, (float)x, (float)y, bc, zi))
								{
									const unsigned long long key = ((unsigned long long)(uint32_t)zi << 32) | prim;
									atomicMin(&sKey[(x - ox) + (y - oy) * kTile], key);
									if (COUNT) ++nFrags;
								}
							}
					}
				}
			__syncthreads();
			if (sSortFallback)
			{
				// too many or interleaved runs for warp 0 (incoherent warps, cooperatively binned big triangles): bitonic network
				sortStagedListFallback(A, item, shard, sIds, n);
				if (tid == 0) atomicAdd(&A.frameInfo[6], 1ull);
			}
			const uint32_t nBigTotal = sNumBig[0];
			if (nBigTotal <= (uint32_t)kListChunk)
				pixelParallel(nBigTotal);
			else
			{
				// Sweep 2 (more large triangles than sBig holds): chunk by chunk, large triangles only - the small ones
				// are done.  Two barriers per chunk: the counter of the NEXT chunk is reset while the current one is being
				// filled, and the trailing barrier protects sBig from the next chunk.
				__syncthreads();                               // everybody has read the total
				if (tid < 2) sNumBig[tid] = 0;
				__syncthreads();
				for (uint32_t base = 0, it2 = 0; base < n; base += kListChunk, ++it2)
				{
					if (tid == 0) sNumBig[(it2 + 1u) & 1u] = 0;
					if (base + (uint32_t)tid < n)
					{
						TriShared t;
						int x0, x1, y0, y1;
						if (expandEntry(idAt(base + tid), t, x0, x1, y0, y1) > kSmallArea)
							packBig(t, (uint32_t)y0 | ((uint32_t)y1 << 16), sBig[atomicAdd(&sNumBig[it2 & 1u], 1u)]);
					}
					__syncthreads();
					pixelParallel(sNumBig[it2 & 1u]);
					if (base + kListChunk < n) __syncthreads();
				}
			}
			best = (bestPrim == kNoPrim) ? kEmptyKey : (((unsigned long long)(uint32_t)bestZ << 32) | bestPrim);
			best = min(best, sKey[tid]);
			if (COUNT)
			{
				unsigned int v = nFrags;
				for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, d);
				if ((tid & 31) == 0 && v) atomicAdd(&A.frameInfo[4], (unsigned long long)v);
			}
		}
		if (desc.w)
		{
			// direct path: the fragments the setup kernel resolved for this block's pixels; the key buffer is left empty for
			// the next frame that uses it.  Sharded front end: desc.w is a mask of ranks - this rank's own key buffer, and the
			// blocks the peers pushed into the inbox (16x16 bins: one 256-key block per owned tile)
			const bool own = !shard || ((desc.w >> fp.ownerRank) & 1u);
			if (own)
			{
				if (live)
				{
					unsigned long long* kp = A.keys + keyIndex(fp.keyBlocksX, px, py);
					const unsigned long long k = *kp;
					if (k != kEmptyKey) { *kp = kEmptyKey; best = min(best, k); }
				}
				if (tid == 0 && sub == 0) A.dirty[bin] = 0;
			}
			if (shard)
			{
				const size_t slot = ((size_t)(bin / fp.ownerCount) << 8) + (size_t)tid;
				for (uint32_t p = 0; p < fp.ownerCount; ++p)
					if (p != fp.ownerRank && ((desc.w >> p) & 1u))
					{
						best = min(best, A.shard.keyInbox[p][slot]);
						if (tid == 0) A.shard.dirtyInbox[p][bin] = 0;
					}
			}
		}
		if (!live) return true;

		// ---- deferred shade of the winner (PS_Kernel, CUDARenderer.cu:833-897) + ROP write-back
		const size_t pix = (size_t)px + (size_t)py * fp.width;
		uint32_t colour;
		int depthOut = 0x7FFFFFFF;
		float zOut = 0.f;
		uint32_t prim = kNoPrim;
		if (best == kEmptyKey)
		{
			// untouched fragment: zero-cleared PixelShade whose colour32 is the clear colour
			if (fp.visState == 0u) colour = packColour(fp.clearRGB[0], fp.clearRGB[1], fp.clearRGB[2]);
			else colour = 0u;   // Depth: (0-.985)/.015*255 < 0 -> 0 ; Normal: n = 0
		}
		else
		{
			prim = (uint32_t)best;
			depthOut = (int)(uint32_t)(best >> 32);
			const uint4 tr = __ldg(A.triRefs + prim);
			float4 p0, p1, p2;
			fetchPositions<OTF>(A, tr, p0, p1, p2);
			TriShared t;
			{
				TriRecord r;
				rasterConstants(screenMap(p0, p1, p2, fp, r), fp, r);
				expandRecord(r, prim, t);
			}
			Bary bc; int zi;
			fragmentTest<false>(t, fx, fy, bc, zi);      // same arithmetic as the coverage pass
			zOut = bc.z;
			const MeshConst& mc = A.meshes[tr.w];
			if (fp.visState == 1u)
			{
				const float r = __fdiv_rn(bc.z - 0.985f, 1.f - 0.985f);   // Remap, MathUtilities.h:175-178
				colour = (__float2uint_rz(r * 255.f) & 0xFFu) << 16;
			}
			else
			{
				Interp S;
				S.b0 = bc.b0; S.b1 = bc.b1; S.b2 = bc.b2;
				S.iw0 = __frcp_rn(p0.w); S.iw1 = __frcp_rn(p1.w); S.iw2 = __frcp_rn(p2.w);
				const uint32_t inDelta = mc.vtxInBase - mc.vtxOutBase;     // transformed-stream index -> IVertex record (mod 2^32)
				const float2* r0 = A.ivertex + (size_t)(tr.x + inDelta) * 7;
				const float2* r1 = A.ivertex + (size_t)(tr.y + inDelta) * 7;
				const float2* r2 = A.ivertex + (size_t)(tr.z + inDelta) * 7;
				const bool textured = TEXTURED && (fp.visState == 0u) && mc.tex[0] >= 0;
				if (!TEXTURED || (fp.visState == 0u && !textured))
				{
					// vertex colour (IVertex floats 11..13, copied unchanged by GetNDCVertex), NOT multiplied
					// by wInterpolated (CUDAROPs.cu:379); read from the colour stream (same values, indexed like the positions)
					const float4 c0 = __ldg(A.colour + tr.x), c1 = __ldg(A.colour + tr.y), c2 = __ldg(A.colour + tr.z);
					colour = packColour(S(c0.x, c1.x, c2.x), S(c0.y, c1.y, c2.y), S(c0.z, c1.z, c2.z));
				}
				else if (fp.visState == 0u && mc.tex[1] < 0)
				{
					// diffuse map only: ShadePixelSafe returns the diffuse sample (CUDAROPs.cu:465-468); only uv is needed
					const float wInterp = __frcp_rn(sum3(S.iw0, bc.b0, S.iw1, bc.b1, S.iw2, bc.b2));
					const float2 a0 = __ldg(r0 + 1), b0 = __ldg(r0 + 2), a1 = __ldg(r1 + 1), b1 = __ldg(r1 + 2), a2 = __ldg(r2 + 1), b2 = __ldg(r2 + 2);
					const float u = S(a0.y, a1.y, a2.y) * wInterp, vv = S(b0.x, b1.x, b2.x) * wInterp;
					const RGB c = sampleTex(A.textures[mc.tex[0]], u, vv, fp.sampleState);
					colour = packColour(c.r, c.g, c.b);
				}
				else
				{
					// transform the three vertices' attributes now (the reference did it for every vertex in VS_Kernel)
					float v[14];
					VertexAttrs t0, t1, t2;
					#pragma unroll
					for (int k = 0; k < 7; ++k) { const float2 q = __ldg(r0 + k); v[2 * k] = q.x; v[2 * k + 1] = q.y; }
					transformAttributes(v, mc, fp.camPos, t0);
					#pragma unroll
					for (int k = 0; k < 7; ++k) { const float2 q = __ldg(r1 + k); v[2 * k] = q.x; v[2 * k + 1] = q.y; }
					transformAttributes(v, mc, fp.camPos, t1);
					#pragma unroll
					for (int k = 0; k < 7; ++k) { const float2 q = __ldg(r2 + k); v[2 * k] = q.x; v[2 * k + 1] = q.y; }
					transformAttributes(v, mc, fp.camPos, t2);
					const float wInterp = __frcp_rn(sum3(S.iw0, bc.b0, S.iw1, bc.b1, S.iw2, bc.b2));
					const float nx = S(t0.n[0], t1.n[0], t2.n[0]) * wInterp, ny = S(t0.n[1], t1.n[1], t2.n[1]) * wInterp, nz = S(t0.n[2], t1.n[2], t2.n[2]) * wInterp;
					if (fp.visState == 2u)
						colour = packColour(nx, ny, nz);
					else
					{
						const float u = S(t0.uv[0], t1.uv[0], t2.uv[0]) * wInterp, vv = S(t0.uv[1], t1.uv[1], t2.uv[1]) * wInterp;
						const float tx = S(t0.tan[0], t1.tan[0], t2.tan[0]), ty = S(t0.tan[1], t1.tan[1], t2.tan[1]), tz = S(t0.tan[2], t1.tan[2], t2.tan[2]);   // no wInterp (CUDAROPs.cu:372)
						float vx = S(t0.vd[0], t1.vd[0], t2.vd[0]), vy = S(t0.vd[1], t1.vd[1], t2.vd[1]), vz = S(t0.vd[2], t1.vd[2], t2.vd[2]);
						normalize3(vx, vy, vz);
						const RGB c = shadePixel(A, mc, fp.sampleState, u, vv, nx, ny, nz, tx, ty, tz, vx, vy, vz);
						colour = packColour(c.r, c.g, c.b);
					}
				}
			}
		}
		storePixel(bin, sub, (uint32_t)tid, pix, colour, depthOut, zOut, prim);
	}
	return true;
}

// The grid of the raster kernel.  The host cannot know how many work items a frame has (the scan kernel counts them on the
// device) and a grid for the worst case - every tile a work item - is mostly CTAs that find nothing and leave: on a 4K
// frame of BASELINE config 3 two thirds of 32 400, ~2 ns each = 40 of the kernel's 117 us (measured with a clamped grid).
// So the library sizes the grid by the item count of the last frame it has seen finish (+ 25 % + 256) and launches a TAIL
// behind it: a few CTAs that stride over whatever items lie beyond the main grid - normally none.  The tail is the same
// code in a loop, compiled with its own register budget (the loop costs the one-item kernel ~100 bytes of spills).
template <bool TOPLEFT, bool COUNT, bool TEXTURED, bool OTF>
__global__ void __launch_bounds__(kTile * kTile, rasterCtasPerSM(TEXTURED, OTF)) raster_shade_kernel(const __grid_constant__ RasterArgs A, const __grid_constant__ FrameParams fp)
{
	pdlPrologue();
	rasterWorkItem<TOPLEFT, COUNT, TEXTURED, OTF>(A, fp, blockIdx.x);
}
template <bool TOPLEFT, bool COUNT, bool TEXTURED, bool OTF>
__global__ void __launch_bounds__(kTile * kTile, 2) raster_tail_kernel(const __grid_constant__ RasterArgs A, const __grid_constant__ FrameParams fp, uint32_t firstItem)
{
	pdlPrologue();
	for (uint32_t item = firstItem + blockIdx.x; rasterWorkItem<TOPLEFT, COUNT, TEXTURED, OTF>(A, fp, item); item += gridDim.x)
		__syncthreads();      // (threads that left the item early must not touch the shared state of the next one yet)
}

// Sharded front end + direct path.  Every rank resolves the small triangles of ITS primitive slice into its own key
// buffer (triangle_setup_kernel), for the whole screen; this kernel then hands the touched blocks to the ranks that own
// the tiles: one warp per tile, a 2 KB block (16x16 keys) travels as 16-byte posted NVLink stores into slot `self` of
// the owner's key inbox, followed by the tile's byte in the owner's map; the local block is reset on the way, so the
// key buffer is empty again for the frame that uses it next.  Blocks of tiles owned here stay where they are (the
// raster kernel merges them like on a single GPU).  A frame of sub-pixel triangles exchanges a few MB per rank
// instead of bin lists.  Needs 16x16 bins on the key buffer's block grid (tile index = block index).
__global__ void __launch_bounds__(256) shard_push_keys_kernel(unsigned long long* __restrict__ keys, uint8_t* __restrict__ dirty,
	const __grid_constant__ ShardKeyPush P, uint32_t nTiles, uint32_t self, uint32_t nRanks)
{
	pdlPrologue();
	// A fixed grid of warps strides over the tiles (a CTA per 8 tiles was 4 050 short-lived CTAs on a 4K frame, ~25 us of
	// CTA turnover for a few MB of traffic).  Warp w looks at tiles w, w + W, w + 2 W, ...: lane l reads the map byte of
	// the l-th of them, so one load instruction covers 32 of the warp's tiles, and the touched ones - neighbours on the
	// screen, hence spread over different warps by the stride - are pushed by the whole warp one after the other.
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t W = gridDim.x * (blockDim.x >> 5), w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	for (uint32_t k0 = 0; (unsigned long long)w + (unsigned long long)k0 * W < nTiles; k0 += 32u)
	{
		const unsigned long long mine = (unsigned long long)w + (unsigned long long)(k0 + lane) * W;
		const bool push = mine < nTiles && (uint32_t)mine % nRanks != self && dirty[mine] != 0;
		unsigned m = __ballot_sync(0xFFFFFFFFu, push);
		while (m)
		{
			const uint32_t bit = (uint32_t)__ffs((int)m) - 1u;
			m &= m - 1u;
			const uint32_t tile = w + (k0 + bit) * W;
			const uint32_t owner = tile % nRanks;
			uint4* src = reinterpret_cast<uint4*>(keys + ((size_t)tile << 8));
			uint4* dst = reinterpret_cast<uint4*>(P.keyInbox[owner] + ((size_t)(tile / nRanks) << 8));
			const uint4 empty = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
			uint4 v[4];
			#pragma unroll
			for (uint32_t i = 0; i < 4; ++i) v[i] = src[i * 32u + lane];
			#pragma unroll
			for (uint32_t i = 0; i < 4; ++i)
			{
				dst[i * 32u + lane] = v[i];
				src[i * 32u + lane] = empty;
			}
			if (lane == 0) { P.dirtyInbox[owner][tile] = 1; dirty[tile] = 0; }
		}
	}
}

// Sharded front end, after the "lists ready" barrier: the work order of this rank's raster kernel.  The length of a
// tile's list is the sum of the heads the ranks pushed for it; the owned tiles are grouped into the same length
// classes tile_scan_kernel uses on a single GPU (longest first), entries {tile, merged length, 0, 0}.  Also raises the
// overflow flag if any rank's list buffer was too small (its marker word rides behind its offsets).
__global__ void __launch_bounds__(256) shard_merge_heads_kernel(const __grid_constant__ ShardTable T, uint32_t nTiles, uint32_t rank, uint32_t nRanks,
	uint4* __restrict__ tileOrder, unsigned long long* __restrict__ frameInfo, const uint8_t* __restrict__ dirty /*direct path: this rank's own map, or null*/)
{
	pdlPrologue();
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
	const uint32_t tile = rank + k * nRanks;
	if (k < nRanks && __ldg(T.tileOffset[k] + nTiles) == 0xFFFFFFFFu) frameInfo[1] = 1ull;
	uint32_t n = 0, cls = kTileClasses, keyMask = 0;
	if (tile < nTiles)
	{
		for (uint32_t p = 0; p < nRanks; ++p) n += __ldg(T.tileCount[p] + tile);
		// direct path: which ranks hold fragments of this tile in a key buffer (bit p: rank p's block in the inbox; bit
		// `rank`: this rank's own key buffer).  Plain loads: the peers' bytes arrived before the "lists ready" barrier.
		if (dirty)
			for (uint32_t p = 0; p < nRanks; ++p)
				if ((p == rank ? dirty[tile] : T.dirtyInbox[p][tile]) != 0) keyMask |= 1u << p;
		cls = (n == 0u && keyMask) ? kTileClasses - 2u : tileClass(n);
	}
	#pragma unroll
	for (uint32_t c = 0; c < kTileClasses; ++c)
	{
		const unsigned m = __ballot_sync(0xFFFFFFFFu, cls == c);
		if (!m) continue;
		uint32_t base = 0;
		if (lane == 0) base = (uint32_t)atomicAdd(&frameInfo[kInfoClassBase + c], (unsigned long long)__popc(m));
		base = __shfl_sync(0xFFFFFFFFu, base, 0);
		if (cls == c) tileOrder[(size_t)c * nTiles + base + (uint32_t)__popc(m & ((1u << lane) - 1u))] = make_uint4(tile, n, 0u, keyMask);
	}
}

// CUDARenderer::DrawTexture / DrawTextureGlobal (testing aids of the reference, CUDARenderer.cu:1565-1620, 1968-2012): the
// texture stretched over the window (or texel-for-pixel), no geometry.  global = 0: TextureTestKernel - the texel CUDA's
// point filter returns for normalised (x / width, y / height), bytes 0 and 2 swapped, alpha kept; global = 1:
// DrawTextureGlobalKernel - CUDATextureSampler::Sample + the truncating colour pack.  The reference launches
// (width / 16) x (height / 16) blocks of 16 x 16 threads: the remainder columns / rows keep the cleared colour.
__global__ void __launch_bounds__(256) draw_texture_kernel(uint32_t* __restrict__ framebuffer, TexDesc tex, uint32_t width, uint32_t height,
	uint32_t global, uint32_t stretched, uint32_t sampleState, uint32_t clearColour)
{
	const uint32_t x = blockIdx.x * 16u + (threadIdx.x & 15u), y = blockIdx.y * 16u + (threadIdx.x >> 4);
	if (x >= width || y >= height) return;
	uint32_t colour = clearColour;
	if (x < (width / 16u) * 16u && y < (height / 16u) * 16u)
	{
		if (!global)
		{
			const float u = __fdiv_rn((float)x, (float)width), v = __fdiv_rn((float)y, (float)height);
			const uint32_t tx = min((uint32_t)__float2int_rd(u * (float)tex.w), tex.w - 1u), ty = min((uint32_t)__float2int_rd(v * (float)tex.h), tex.h - 1u);
			const uint32_t s = __ldg(tex.px + texelIndex(tex, tx + ty * tex.w));
			colour = (s & 0xFF00FF00u) | ((s & 0xFFu) << 16) | ((s >> 16) & 0xFFu);
		}
		else
		{
			const float u = __fdiv_rn((float)x, stretched ? (float)width : (float)tex.w), v = __fdiv_rn((float)y, stretched ? (float)height : (float)tex.h);
			const RGB c = sampleTex(tex, u, v, sampleState);
			colour = packColour(c.r, c.g, c.b);
		}
	}
	framebuffer[(size_t)x + (size_t)y * width] = colour;
}

// Distributed Present: copy the 16x16 tiles this rank owns from its colour surface into the row-major frame in mapped
// host memory (zero-copy stores over this GPU's own PCIe link).  64 threads per tile: one 16-byte store per thread, a
// tile row = one 64-byte segment.  Four tiles per CTA.
__global__ void __launch_bounds__(256) present_tiles_kernel(const uint32_t* __restrict__ colour, uint32_t* __restrict__ hostFrame,
	uint32_t rank, uint32_t count, uint32_t owned, FrameParams fp)
{
	const uint32_t k = blockIdx.x * 4u + (threadIdx.x >> 6), t = threadIdx.x & 63u;
	if (k >= owned * fp.subX * fp.subY) return;
	uint32_t sub, binX, subX;
	const uint32_t slot = divMagic(k, fp.subX * fp.subY, fp.subsMagic, sub);
	const uint32_t bin = rank + slot * count;
	const uint32_t binY = divMagic(bin, fp.binsX, fp.binsXMagic, binX);
	const uint32_t subY = divMagic(sub, fp.subX, fp.subXMagic, subX);
	const int ox = (int)(binX * fp.binW + subX * kTile), oy = (int)(binY * fp.binH + subY * kTile);
	const int ex = min(min(ox + kTile, (int)((binX + 1) * fp.binW)), (int)fp.width);
	const int ey = min(min(oy + kTile, (int)((binY + 1) * fp.binH)), (int)fp.height);
	const int py = oy + (int)(t >> 2), px = ox + (int)(t & 3u) * 4;
	if (py >= ey || px >= ex) return;
	const size_t pix = (size_t)px + (size_t)py * fp.width;
	if (px + 4 <= ex && (pix & 3u) == 0u)
		*reinterpret_cast<uint4*>(hostFrame + pix) = *reinterpret_cast<const uint4*>(colour + pix);
	else
		for (int x = px; x < min(px + 4, ex); ++x) hostFrame[(size_t)x + (size_t)py * fp.width] = colour[(size_t)x + (size_t)py * fp.width];
}

// sort-first composition on the gathering rank: scatter one rank's packed tile-major colours back
// into the row-major framebuffer (tile owner = tile index mod count)
__global__ void __launch_bounds__(kTile * kTile) unpack_tiles_kernel(const uint32_t* __restrict__ packed, uint32_t* __restrict__ framebuffer,
	uint32_t rank, uint32_t count, FrameParams fp)
{
	const uint32_t bin = rank + blockIdx.x * count;
	const uint32_t binX = bin % fp.binsX, binY = bin / fp.binsX;
	const uint32_t subX = blockIdx.y % fp.subX, subY = blockIdx.y / fp.subX;
	const int tid = threadIdx.x;
	const int ox = (int)(binX * fp.binW + subX * kTile), oy = (int)(binY * fp.binH + subY * kTile);
	const int ex = min(min(ox + kTile, (int)((binX + 1) * fp.binW)), (int)fp.width);
	const int ey = min(min(oy + kTile, (int)((binY + 1) * fp.binH)), (int)fp.height);
	const int px = ox + (tid & (kTile - 1)), py = oy + (tid >> 4);
	if (px < ex && py < ey)
		framebuffer[(size_t)px + (size_t)py * fp.width] = packed[((size_t)blockIdx.x * gridDim.y + blockIdx.y) * (kTile * kTile) + tid];
}

}
