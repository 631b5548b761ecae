// Geometry front end: vertex transform, triangle setup + cull, tile binning
// (count -> scan -> scatter -> sort).  Replaces VS_Kernel, TA_Kernel, TB_Kernel and
// CUDAAtomicQueue of the reference (CUDARenderer.cu:673-796, 903-988; CUDAAtomicQueue.cuh).
#pragma once
#include "device_types.cuh"

namespace b2r {

// a*b + c*d + e*f the way the reference's device code evaluates it (SASS: FMUL c*d, FFMA a*b, FFMA e*f)
__device__ __forceinline__ float sum3(float a, float b, float c, float d, float e, float f)
{
	return fmaf(e, f, fmaf(a, b, c * d));
}
// a*b - c*d  (SASS: FMUL c*d, FFMA a*b - t)
__device__ __forceinline__ float diff2(float a, float b, float c, float d)
{
	return fmaf(a, b, -(c * d));
}
// Matrix4 * point (w = 1), row r of a data[col][row] matrix: Matrix4.h:158-166
__device__ __forceinline__ float mulRow4(const float* m, int r, float x, float y, float z)
{
	return sum3(m[0 + r], x, m[4 + r], y, m[8 + r], z) + m[12 + r];
}
__device__ __forceinline__ float mulRow3(const float* m, int r, float x, float y, float z)
{
	return sum3(m[0 + r], x, m[3 + r], y, m[6 + r], z);
}
// Vector.h:42-56 Normalize: v * (1 / sqrt(dot)), zero vector when the length is 0
__device__ __forceinline__ void normalize3(float& x, float& y, float& z)
{
	const float m = __fsqrt_rn(sum3(x, x, y, y, z, z));
	if (m == 0.f) { x = 0.f; y = 0.f; z = 0.f; return; }
	const float inv = __frcp_rn(m);
	x = x * inv; y = y * inv; z = z * inv;
}

// Back-face test of TA_Kernel (CUDARenderer.cu:760-777): d = dot(normalize(n), camFwd), the triangle is culled when
// d <= 0.  Only the SIGN of d is used, so the IEEE square root and reciprocal of the normalisation are skipped whenever
// the sign is certain without them, and the reference's arithmetic (`exact` below) runs only for the rest.
//
// Why the short cut returns the same answer.  Write u = 2^-24, L = |n|, S = sum n_i c_i, T = sum |n_i c_i|,
// M = sum |n_i|.  Under the guard (kCullFastOk: every |c_i| <= 2^20, set by the host; 2^-40 < M < 2^40):
//  * the squared length lies in (2^-84, 2^82) - no overflow, squares that underflow change it by < 2^-65 relative - so the
//    reference's m = sqrt_rn(len2) and inv = rcp_rn(m) carry relative errors <= 2u and <= 3.5u, and its normalised
//    components are n_i / L * (1 + e_i), |e_i| <= 4.5u, or within 2^-150 of that when the product is subnormal;
//  * its d is three roundings away from sum nhat_i c_i, so d_ref = (1 / L) * sum n_i c_i (1 + k_i) + E with |k_i| <= 8u
//    and |E| <= 2^-127 (underflows, times |c_i| <= 2^20);
//  * the S computed here differs from the exact S by <= 3u T + 2^-147.
// With |S_computed| > 64u T (the 2^-18 below) the exact S has the sign of the computed one and |S| > 60u T, so the
// reference's numerator sum n_i c_i (1 + k_i) has that sign too and magnitude > 52u T; and |S| >= 2^-60 M together with
// L <= M puts |d_ref| above 2^-61, far above E and the subnormal range: d_ref is a non-zero number of the sign of S.
// Anything else - zero-length or huge normals, NaN / Inf operands (every comparison below is false for them), a dot
// product within rounding of 0 - takes the exact path.  tests/test_gpu_units.py::test_sign_only_cull_matches_the_exact_test
// compares the two on 2^32 adversarial triangles per seed.
constexpr uint32_t kCullFastOk = 0x100u;      // bit of FrameParams::cullMode
#ifndef B2R_CULL_MARGIN
#define B2R_CULL_MARGIN 3.814697265625e-6f      // 2^-18 = 64u (a build with 2^-25 fails the self test: the test does reach the boundary)
#endif
__device__ __forceinline__ bool backFaceExact(float nx, float ny, float nz, const float* c, uint32_t cull)
{
	normalize3(nx, ny, nz);
	// the two culling branches of TA_Kernel were contracted differently (SASS @0x06f0 / @0x0aa0)
	const float d = (cull == 0u) ? fmaf(nz, c[2], fmaf(nx, c[0], ny * c[1]))
	                             : fmaf(nz, c[2], fmaf(ny, c[1], nx * c[0]));
	return d <= 0.f;                // the live FrontFace branch uses the same test (CUDARenderer.cu:773)
}
// the short cut: true when the sign of the reference's d is certain (then `back` is the answer)
__device__ __forceinline__ bool backFaceCertain(float nx, float ny, float nz, const float* c, bool& back)
{
	const float S = fmaf(nz, c[2], fmaf(nx, c[0], ny * c[1]));
	const float T = fmaf(fabsf(nz), fabsf(c[2]), fmaf(fabsf(nx), fabsf(c[0]), fabsf(ny) * fabsf(c[1])));
	const float M = fabsf(nx) + fabsf(ny) + fabsf(nz);
	const float aS = fabsf(S);
	back = S < 0.f;
	return aS > T * B2R_CULL_MARGIN && aS >= M * 8.673617379884035e-19f /*2^-60*/ &&
		M > 9.094947017729282e-13f /*2^-40*/ && M < 1.099511627776e12f /*2^40*/;
}
__device__ __forceinline__ bool backFace(float nx, float ny, float nz, const float* c, uint32_t cullMode)
{
#ifndef B2R_EXACT_CULL_ONLY
	bool back;
	if ((cullMode & kCullFastOk) != 0u && backFaceCertain(nx, ny, nz, c, back)) return back;
#endif
	return backFaceExact(nx, ny, nz, c, cullMode & 0xFFu);
}

// ---------------------------------------------------------------------------------------
// upload-time: pull the positions out of the 56 B AoS records into a 16 B stream, so the per-frame
// transform reads 16 B per vertex instead of dragging the whole record through the memory system
__global__ void __launch_bounds__(256) split_positions_kernel(const float2* __restrict__ ivertex, float4* __restrict__ posIn, uint32_t first, uint32_t n)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float2* rec = ivertex + (size_t)(first + i) * 7;
	const float2 a = __ldg(rec), b = __ldg(rec + 1);
	posIn[first + i] = make_float4(a.x, a.y, b.x, 1.f);
}

// clip = WVP * (p, 1); ndc = clip.xyz / clip.w (IEEE); w kept.  GetNDCVertex, CUDAROPs.cu:99-107.
// One function for every place a position is transformed (the vertex stage, and the raster kernel when it
// transforms the vertices of a list entry itself - sharded front end with a sliced vertex stage), hence the same bits.
__device__ __forceinline__ float4 transformPosition(const float* m /*WVP, data[col][row]*/, float x, float y, float z)
{
	const float cx = mulRow4(m, 0, x, y, z);
	const float cy = mulRow4(m, 1, x, y, z);
	const float cz = mulRow4(m, 2, x, y, z);
	const float cw = mulRow4(m, 3, x, y, z);
	float4 p;
	p.x = __fdiv_rn(cx, cw); p.y = __fdiv_rn(cy, cw); p.z = __fdiv_rn(cz, cw); p.w = cw;
	return p;
}

// vertex transform: one CTA per kVsPerCta chunks of 256 vertices (a chunk belongs to one mesh); 16 B in, 16 B out, coalesced.
// The kernel is latency-bound (chunk descriptor -> position -> 3 IEEE divisions -> store), so a thread works on kVsPerCta
// vertices whose descriptor and position loads are all issued before the first is used.
#ifndef B2R_VS_PER_CTA
#define B2R_VS_PER_CTA 4
#endif
constexpr uint32_t kVsPerCta = B2R_VS_PER_CTA;
__global__ void __launch_bounds__(kChunk) vertex_transform_kernel(
	const Chunk* __restrict__ chunks, uint32_t nChunks, const MeshConst* __restrict__ meshes,
	const float4* __restrict__ posIn, float4* __restrict__ pos4)
{
	pdlPrologue();
	__shared__ float sMat[kVsPerCta][16];
	uint32_t mesh[kVsPerCta], count[kVsPerCta], inStart[kVsPerCta], outStart[kVsPerCta];
	float4 in[kVsPerCta];
	#pragma unroll
	for (uint32_t k = 0; k < kVsPerCta; ++k)
	{
		const uint32_t c = blockIdx.x * kVsPerCta + k;
		uint2 a = make_uint2(0u, 0u), b = make_uint2(0u, 0u);
		if (c < nChunks)
		{
			const uint2* d = reinterpret_cast<const uint2*>(chunks + c);      // {mesh, count}, {inStart, outStart}, ...
			a = __ldg(d); b = __ldg(d + 1);
		}
		mesh[k] = a.x; count[k] = a.y; inStart[k] = b.x; outStart[k] = b.y;
	}
	#pragma unroll
	for (uint32_t k = 0; k < kVsPerCta; ++k)
	{
		in[k] = make_float4(0.f, 0.f, 0.f, 1.f);
		if (threadIdx.x < count[k]) in[k] = __ldcs(posIn + (size_t)inStart[k] + threadIdx.x);
	}
	#pragma unroll
	for (uint32_t k = 0; k < kVsPerCta; ++k)
		if ((threadIdx.x >> 4) == k && count[k]) sMat[k][threadIdx.x & 15u] = meshes[mesh[k]].wvp[threadIdx.x & 15u];
	__syncthreads();
	#pragma unroll
	for (uint32_t k = 0; k < kVsPerCta; ++k)
		if (threadIdx.x < count[k]) pos4[(size_t)outStart[k] + threadIdx.x] = transformPosition(sMat[k], in[k].x, in[k].y, in[k].z);
}

// sharded front end with a sliced vertex stage: the range of transformed-vertex indices the primitives
// [primLo, primHi) refer to (static: evaluated once per scene and slice).  out[0] = min (preset to 0xFFFFFFFF), out[1] = max (preset to 0)
__global__ void __launch_bounds__(256) prim_vertex_range_kernel(const uint4* __restrict__ triRefs, uint32_t primLo, uint32_t primHi, uint32_t* __restrict__ out)
{
	uint32_t lo = 0xFFFFFFFFu, hi = 0u;
	for (uint32_t p = primLo + blockIdx.x * blockDim.x + threadIdx.x; p < primHi; p += gridDim.x * blockDim.x)
	{
		const uint4 t = __ldg(triRefs + p);
		lo = min(lo, min(t.x, min(t.y, t.z)));
		hi = max(hi, max(t.x, max(t.y, t.z)));
	}
	lo = __reduce_min_sync(0xFFFFFFFFu, lo); hi = __reduce_max_sync(0xFFFFFFFFu, hi);
	if ((threadIdx.x & 31u) == 0u && lo <= hi) { atomicMin(out, lo); atomicMax(out + 1, hi); }
}

// the attribute half of GetNDCVertex (CUDAROPs.cu:108-115) for ONE vertex record
struct VertexAttrs { float uv[2], n[3], tan[3], vd[3], c[3]; };
__device__ __forceinline__ void transformAttributes(const float* v /*14 floats of the IVertex*/, const MeshConst& mc, const float* camPos, VertexAttrs& o)
{
	o.uv[0] = v[3]; o.uv[1] = v[4];
	o.n[0] = mulRow3(mc.rot, 0, v[5], v[6], v[7]); o.n[1] = mulRow3(mc.rot, 1, v[5], v[6], v[7]); o.n[2] = mulRow3(mc.rot, 2, v[5], v[6], v[7]);
	o.tan[0] = mulRow3(mc.rot, 0, v[8], v[9], v[10]); o.tan[1] = mulRow3(mc.rot, 1, v[8], v[9], v[10]); o.tan[2] = mulRow3(mc.rot, 2, v[8], v[9], v[10]);
	float vx = mulRow4(mc.world, 0, v[0], v[1], v[2]) - camPos[0];
	float vy = mulRow4(mc.world, 1, v[0], v[1], v[2]) - camPos[1];
	float vz = mulRow4(mc.world, 2, v[0], v[1], v[2]) - camPos[2];
	normalize3(vx, vy, vz);
	o.vd[0] = vx; o.vd[1] = vy; o.vd[2] = vz;
	o.c[0] = v[11]; o.c[1] = v[12]; o.c[2] = v[13];
}

// parity surface: the full 72 B OVertex records of one mesh, produced on demand (b2r_read_overtex)
__global__ void __launch_bounds__(256) overtex_kernel(const MeshConst* __restrict__ meshes, uint32_t mesh,
	const float2* __restrict__ ivertex, const float4* __restrict__ pos4, float* __restrict__ out18, FrameParams fp)
{
	const MeshConst& mc = meshes[mesh];
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= mc.nVertices) return;
	float v[14];
	const float2* rec = ivertex + (size_t)(mc.vtxInBase + i) * 7;
	#pragma unroll
	for (int k = 0; k < 7; ++k) { const float2 t = __ldg(rec + k); v[2 * k] = t.x; v[2 * k + 1] = t.y; }
	VertexAttrs a;
	transformAttributes(v, mc, fp.camPos, a);
	const float4 p = pos4[mc.vtxOutBase + i];
	float* o = out18 + (size_t)i * 18;
	o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = p.w;
	o[4] = a.uv[0]; o[5] = a.uv[1];
	o[6] = a.n[0]; o[7] = a.n[1]; o[8] = a.n[2];
	o[9] = a.tan[0]; o[10] = a.tan[1]; o[11] = a.tan[2];
	o[12] = a.vd[0]; o[13] = a.vd[1]; o[14] = a.vd[2];
	o[15] = a.c[0]; o[16] = a.c[1]; o[17] = a.c[2];
}

// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float min3ref(float a, float b, float c) { float m = a; if (b < m) m = b; if (c < m) m = c; return m; }
__device__ __forceinline__ float max3ref(float a, float b, float c) { float m = a; if (b > m) m = b; if (c > m) m = c; return m; }
__device__ __forceinline__ int clampi(int v, int lo, int hi) { const int c = v < lo ? lo : v; return c > hi ? hi : c; }
// (short)f -+ 1 then ClampFast to [0, limit] as shorts: CUDAROPs.cu:184-192
__device__ __forceinline__ int bboxCoord(float f, int delta, int limit)
{
	const int t = __float2int_rz(f) + delta;          // cvt.rzi.s32.f32, saturating, NaN -> 0
	const int s = (int)(short)(t & 0xFFFF);           // wraps exactly like the 16-bit store
	return clampi(s, 0, (int)(short)limit);
}
__device__ __forceinline__ bool tileOwned(const FrameParams& fp, uint32_t tile)
{
	return fp.ownerCount <= 1 || fp.shardGeometry || (tile % fp.ownerCount) == fp.ownerRank;
}


// Warp-aggregated per-tile counter update: the lanes of a warp that address the same tile in the
// same step are combined into ONE atomic (consecutive triangles of a mesh usually land in the
// same few tiles).  Every lane of the warp must call this the same number of times (`steps` =
// warp-wide maximum of nTiles); lanes with k >= nTiles pass tile = kNoTile.
constexpr uint32_t kNoTile = 0xFFFFFFFFu;
constexpr uint32_t kBinSlotsWord = 2;              // counters[2]: direct path - 32-entry slots of primitives with a small rectangle that went into the lists (what bin_scatter_kernel has to visit)
constexpr uint32_t kRectDirect = 0x80000000u;      // rects[prim].x bit 31: the setup kernel resolved the triangle itself (direct path), it is in no list
// counters[]: [0] visible triangles (summed by tile_scan_kernel), [1] big triangles,
// [kVisBase + 32*s] partial visible counts, one 128-byte line per slot,
// [kScanBase ...] the 64-bit slice sums of the chained tile scan
constexpr uint32_t kVisSlots = 64, kVisBase = 32, kScanBase = kVisBase + kVisSlots * 32 /* 8-byte aligned */, kScanMaxCtas = 2048 /* 2048 x 2048 tiles = a 32767^2 screen */, kCounterWords = kScanBase + 2 * kScanMaxCtas;
template <bool RETURN>
__device__ __forceinline__ uint32_t warpTileAdd(uint32_t* counters, uint32_t tile)
{
	const unsigned peers = __match_any_sync(0xFFFFFFFFu, tile);
	const int leader = __ffs(peers) - 1;
	const unsigned lane = threadIdx.x & 31u;
	uint32_t base = 0;
	if (tile != kNoTile && (int)lane == leader)
	{
		if (RETURN) base = atomicAdd(&counters[tile], (uint32_t)__popc(peers));
		else atomicAdd(&counters[tile], (uint32_t)__popc(peers));
	}
	if (RETURN)
	{
		base = __shfl_sync(0xFFFFFFFFu, base, leader);
		return base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
	}
	return 0;
}

// Per-mesh constants of a frame, built on the device: WVP = viewProj * world in the element order of
// Matrix4.h:120-143 (what the reference computes on the host, CUDARenderer.cu:242; unfused multiplies and adds,
// left to right - this translation unit is compiled with -fmad=false, so the bits equal a host evaluation),
// rot = (FMatrix3)world.  viewProj is read straight from the frame's pinned host slot (zero-copy): the host's
// per-frame work and the PCIe traffic do not depend on the number of meshes.
__device__ __forceinline__ void buildMeshConst(const float* __restrict__ vp, const MeshStatic& ms, MeshConst& mc)
{
	#pragma unroll
	for (int col = 0; col < 4; ++col)
		#pragma unroll
		for (int r = 0; r < 4; ++r)
			mc.wvp[col * 4 + r] = vp[0 * 4 + r] * ms.world[col * 4 + 0] + vp[1 * 4 + r] * ms.world[col * 4 + 1] +
			                      vp[2 * 4 + r] * ms.world[col * 4 + 2] + vp[3 * 4 + r] * ms.world[col * 4 + 3];
	#pragma unroll
	for (int i = 0; i < 16; ++i) mc.world[i] = ms.world[i];
	#pragma unroll
	for (int col = 0; col < 3; ++col)
		#pragma unroll
		for (int r = 0; r < 3; ++r) mc.rot[col * 3 + r] = ms.world[col * 4 + r];
	mc.vtxInBase = ms.vtxInBase; mc.vtxOutBase = ms.vtxOutBase; mc.nVertices = ms.nVertices;
	mc.idxBase = ms.idxBase; mc.nTris = ms.nTris; mc.primBase = ms.primBase;
	mc.tex[0] = ms.tex[0]; mc.tex[1] = ms.tex[1]; mc.tex[2] = ms.tex[2]; mc.tex[3] = ms.tex[3];
	mc.pad = 0;
}
__device__ __forceinline__ void buildMeshConsts(uint32_t gtid, uint32_t gsize, const float* __restrict__ hostViewProj,
	const MeshStatic* __restrict__ meshStatic, MeshConst* __restrict__ meshConst, uint32_t nMeshes)
{
	if (gtid >= nMeshes) return;
	float vp[16];
	#pragma unroll
	for (int i = 0; i < 16; ++i) vp[i] = hostViewProj[i];
	for (uint32_t m = gtid; m < nMeshes; m += gsize) buildMeshConst(vp, meshStatic[m], meshConst[m]);
}

__device__ __forceinline__ unsigned long long globalTimerNs()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
// one warp: spin until the flag words [slotBase, slotBase + n) of this rank have reached `epoch` (stored by the peers over
// NVLink); a peer that never arrives raises *timeoutFlag after ~4 s instead of hanging the GPU
__device__ __forceinline__ void peerFlagsWait(PeerFlags pf, uint32_t self, uint32_t n, uint32_t slotBase, uint32_t epoch, unsigned long long* timeoutFlag)
{
	const uint32_t lane = threadIdx.x & 31u;
	if (lane < n)
	{
		const volatile uint32_t* f = reinterpret_cast<const volatile uint32_t*>(pf.flags[self] + slotBase + lane);
		const unsigned long long t0 = globalTimerNs();
		while ((int)(*f - epoch) < 0)
		{
			if (globalTimerNs() - t0 > 4000000000ull) { *timeoutFlag = 1ull; break; }
			__nanosleep(64);
		}
	}
	__threadfence_system();
}

// start of frame: zero the tile counts (tileCount is padded to a multiple of 4 words) and the counters, and
// build the per-mesh constants.  Together with publish_info_kernel this keeps the frame off the copy engines,
// which belong to Present: a cudaMemcpyAsync here can be queued behind the previous frame's 33 MB framebuffer copy.
__global__ void __launch_bounds__(256) frame_clear_kernel(uint32_t* __restrict__ tileCount, uint32_t nTiles,
	uint32_t* __restrict__ counters, unsigned long long* __restrict__ frameInfo, uint32_t infoWords, uint32_t counterWords,
	const float* __restrict__ hostViewProj, const MeshStatic* __restrict__ meshStatic, MeshConst* __restrict__ meshConst, uint32_t nMeshes,
	const __grid_constant__ PeerFlags waitPf, uint32_t waitSelf, uint32_t waitRanks /*0: no wait*/, uint32_t waitEpoch)
{
	pdlPrologue();
	const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
	buildMeshConsts(gtid, gsize, hostViewProj, meshStatic, meshConst, nMeshes);
	const uint32_t q = (nTiles + 3) / 4;
	for (uint32_t i = gtid; i < q; i += gsize)
		reinterpret_cast<uint4*>(tileCount)[i] = make_uint4(0, 0, 0, 0);
	if (blockIdx.x == 0)
	{
		for (uint32_t i = threadIdx.x; i < counterWords; i += blockDim.x) counters[i] = 0;
		if (threadIdx.x < infoWords) frameInfo[threadIdx.x] = 0;
		// frame ring + sharded front end: the peers must have finished the frame that used this ring slot before (their
		// raster kernels read the list set this frame is about to store into: "frame done" words 16..23); normally long
		// satisfied.  The kernels that store into peer memory follow this one in the stream.
		if (waitRanks && threadIdx.x < 32)
		{
			__syncwarp();
			peerFlagsWait(waitPf, waitSelf, waitRanks, 16u, waitEpoch, frameInfo + 7);
		}
	}
}
// the constants alone, for frames whose vertex stage runs ahead on its own stream
__global__ void __launch_bounds__(64) build_constants_kernel(const float* __restrict__ hostViewProj, const MeshStatic* __restrict__ meshStatic,
	MeshConst* __restrict__ meshConst, uint32_t nMeshes)
{
	pdlPrologue();
	buildMeshConsts(blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, hostViewProj, meshStatic, meshConst, nMeshes);
}
// end of frame: the frame's counters go to the pinned host slot (zero-copy write).  Sharded front end: the same
// kernel tells rank 0 that this rank's colours of the frame are in rank 0's surface (the raster kernel before it
// in the stream has completed, so its NVLink stores have landed).
// Frame ring + sharded front end: it also tells EVERY rank that this rank has finished the frame (flag words 16..23),
// i.e. no longer reads the list set of the frame's ring slot - what a peer waits for before it stores the lists of a
// later frame into that set (doneRanks = 0: no such signal).
__global__ void __launch_bounds__(32) publish_info_kernel(const unsigned long long* __restrict__ info, unsigned long long* __restrict__ hostInfo, uint32_t words,
	uint32_t* landedFlag, uint32_t epoch, const __grid_constant__ PeerFlags donePf, uint32_t doneRanks, uint32_t self)
{
	pdlPrologue();
	if (threadIdx.x < words) hostInfo[threadIdx.x] = info[threadIdx.x];      // visible to the host once the frame's `done` event has completed
	if ((landedFlag || doneRanks) && threadIdx.x == 0) __threadfence_system();
	__syncwarp();
	if (landedFlag && threadIdx.x == 0)
		*reinterpret_cast<volatile uint32_t*>(landedFlag) = epoch;
	if (threadIdx.x < doneRanks)
		*reinterpret_cast<volatile uint32_t*>(donePf.flags[threadIdx.x] + 16 + self) = epoch;
}

// Warp-level binning of up to 32 small rectangles (one per lane, rectangle = inclusive tile range,
// nTiles = 0 for lanes without one).  Consecutive primitives overlap the same few tiles, so the warp
// walks the UNION of its rectangles tile by tile: for each tile the lanes whose rectangle contains it
// are found with one ballot and served by ONE atomic.  SCATTER: the lanes' ids are written in lane
// order, so every (warp, tile) pair contributes exactly one ascending run of primitive ids, and runs
// of different warps cover disjoint id ranges - the property bin_sort_runs_kernel relies on.
// Warps whose union is large (incoherent geometry) fall back to per-lane iteration.
constexpr uint32_t kUnionTilesMax = 64;
// Sharded front end: an id scattered into the list of a tile another rank owns is also stored into the region reserved for
// this rank in the owner's arena, at the same offset as locally (posted NVLink store; the lanes of a run write consecutive
// words, which travel as one transaction), so the exchange of the lists needs no pass of its own.
struct ListMirror { const ShardRegions* regions; uint32_t nRanks, self; };       // nRanks = 0: no mirror
__device__ __forceinline__ void mirrorStore(const ListMirror& M, uint32_t tile, uint32_t pos, uint32_t prim)
{
	if (M.nRanks)
	{
		const uint32_t owner = tile % M.nRanks;
		if (owner != M.self) M.regions->region[owner][pos] = prim;
	}
}

template <bool SCATTER>
__device__ __forceinline__ void warpBinRects(uint32_t nTiles, int bx0, int bx1, int by0, int by1, uint32_t prim,
	uint32_t* __restrict__ counters /*tileCount or tileCursor*/, const uint32_t* __restrict__ tileOffset, uint32_t* __restrict__ binIds,
	const FrameParams& fp, const ListMirror& M = ListMirror{ nullptr, 0u, 0u })
{
	const unsigned lane = threadIdx.x & 31u;
	const unsigned have = __ballot_sync(0xFFFFFFFFu, nTiles > 0);
	if (!have) return;
	const int ux0 = __reduce_min_sync(0xFFFFFFFFu, nTiles ? bx0 : 0x7FFF), ux1 = __reduce_max_sync(0xFFFFFFFFu, nTiles ? bx1 : -1);
	const int uy0 = __reduce_min_sync(0xFFFFFFFFu, nTiles ? by0 : 0x7FFF), uy1 = __reduce_max_sync(0xFFFFFFFFu, nTiles ? by1 : -1);
	const uint32_t uw = (uint32_t)(ux1 - ux0 + 1), uArea = uw * (uint32_t)(uy1 - uy0 + 1);
	if (uArea <= kUnionTilesMax)
	{
		// 32 union tiles at a time: lane u keeps the ballot of tile u, then ALL tiles' atomics are issued by one
		// instruction (one round trip to L2 instead of one per tile), then the ids are written tile by tile
		int tx = ux0, ty = uy0;        // row-major walk over the union, carried across the 32-tile chunks (no div / mod)
		for (uint32_t u0 = 0; u0 < uArea; u0 += 32)
		{
			const uint32_t uEnd = min(uArea - u0, 32u);
			uint32_t myMask = 0, myTile = 0;
			for (uint32_t u = 0; u < uEnd; ++u)
			{
				const bool in = nTiles && tx >= bx0 && tx <= bx1 && ty >= by0 && ty <= by1;
				const unsigned mask = __ballot_sync(0xFFFFFFFFu, in);
				if (lane == u)
				{
					myTile = (uint32_t)tx + (uint32_t)ty * fp.binsX;
					myMask = tileOwned(fp, myTile) ? mask : 0u;
				}
				if (++tx > ux1) { tx = ux0; ++ty; }
			}
			if (SCATTER)
			{
				uint32_t myBase = 0;
				if (myMask) myBase = __ldg(tileOffset + myTile) + atomicAdd(&counters[myTile], (uint32_t)__popc(myMask));
				for (uint32_t u = 0; u < uEnd; ++u)
				{
					const unsigned mask = __shfl_sync(0xFFFFFFFFu, myMask, u);
					const uint32_t base = __shfl_sync(0xFFFFFFFFu, myBase, u);
					const uint32_t tileU = M.nRanks ? __shfl_sync(0xFFFFFFFFu, myTile, u) : 0u;
					if ((mask >> lane) & 1u)
					{
						const uint32_t pos = base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
						binIds[pos] = prim;
						mirrorStore(M, tileU, pos, prim);
					}
				}
			}
			else if (myMask)
				atomicAdd(&counters[myTile], (uint32_t)__popc(myMask));
		}
	}
	else
	{
		const int w = bx1 - bx0 + 1;
		const uint32_t steps = __reduce_max_sync(0xFFFFFFFFu, nTiles);
		for (uint32_t k = 0; k < steps; ++k)
		{
			uint32_t tile = kNoTile;
			if (k < nTiles)
			{
				tile = (uint32_t)(bx0 + (int)(k % (uint32_t)w)) + (uint32_t)(by0 + (int)(k / (uint32_t)w)) * fp.binsX;
				if (!tileOwned(fp, tile)) tile = kNoTile;
			}
			const uint32_t slot = warpTileAdd<SCATTER>(counters, tile);
			if (SCATTER && tile != kNoTile)
			{
				const uint32_t pos = __ldg(tileOffset + tile) + slot;
				binIds[pos] = prim;
				mirrorStore(M, tile, pos, prim);
			}
		}
	}
}

// scene-load time: the static per-primitive reference {v0, v1, v2, mesh} with GLOBAL indices into the
// transformed-vertex stream (mesh.vtxOutBase + index).  It never changes between frames, so the frame
// itself writes nothing per primitive except the 8-byte bin rectangle: the raster kernel rebuilds the
// screen-space triangle of a list entry from this reference and the three transformed positions.
__global__ void __launch_bounds__(kChunk) build_trirefs_kernel(const Chunk* __restrict__ chunks, const uint32_t* __restrict__ indices,
	uint4* __restrict__ triRefs)
{
	const Chunk ch = chunks[blockIdx.x];
	if (threadIdx.x >= ch.count) return;
	const uint32_t* ip = indices + (size_t)ch.inStart + (size_t)threadIdx.x * 3;
	triRefs[ch.primStart + threadIdx.x] = make_uint4(ch.outStart + __ldg(ip), ch.outStart + __ldg(ip + 1), ch.outStart + __ldg(ip + 2), ch.mesh);
}

// scene-load time: the vertex colours (IVertex floats 11..13, which GetNDCVertex copies unchanged) as a 16-byte stream
// indexed like the transformed positions, so the deferred shade of an untextured pixel is vertex reference -> {3 positions,
// 3 colours} in ONE round of independent 16-byte loads instead of reference -> mesh constants -> six 8-byte loads out of
// the 56-byte records
__global__ void __launch_bounds__(kChunk) build_colour_stream_kernel(const Chunk* __restrict__ chunks, const float2* __restrict__ ivertex, float4* __restrict__ colourOut)
{
	const Chunk ch = chunks[blockIdx.x];
	if (threadIdx.x >= ch.count) return;
	const float2* rec = ivertex + (size_t)(ch.inStart + threadIdx.x) * 7;
	const float2 a = __ldg(rec + 5), b = __ldg(rec + 6);
	colourOut[(size_t)ch.outStart + threadIdx.x] = make_float4(a.y, b.x, b.y, 0.f);
}

// screen mapping of the three NDC vertices (CUDAROPs.cu:88-97) and the float bounds
struct ScreenBounds { float minx, maxx, miny, maxy; };
__device__ __forceinline__ ScreenBounds screenMap(const float4& p0, const float4& p1, const float4& p2, const FrameParams& fp, TriRecord& r)
{
	const float W = (float)fp.width, H = (float)fp.height;
	r.x0 = ((p0.x + 1.f) * 0.5f) * W; r.y0 = ((1.f - p0.y) * 0.5f) * H;
	r.x1 = ((p1.x + 1.f) * 0.5f) * W; r.y1 = ((1.f - p1.y) * 0.5f) * H;
	r.x2 = ((p2.x + 1.f) * 0.5f) * W; r.y2 = ((1.f - p2.y) * 0.5f) * H;
	r.z0 = p0.z; r.z1 = p1.z; r.z2 = p2.z;
	ScreenBounds b;
	b.minx = min3ref(r.x0, r.x1, r.x2); b.maxx = max3ref(r.x0, r.x1, r.x2);
	b.miny = min3ref(r.y0, r.y1, r.y2); b.maxy = max3ref(r.y0, r.y1, r.y2);
	return b;
}

// hoisted division only for quotients in [2^-64, ~1] with a sane divisor (else __fdiv_rn): the smallest weight that takes it
__device__ __forceinline__ float divisionThreshold(float area)
{
	return (area >= kDivLo && area <= kDivHi) ? fmaxf(kDivLo, area * 5.42101086e-20f) : __int_as_float(0x7F800000);
}

// the per-triangle constants of the raster kernel's fragment test (area, candidate box, hoisted
// reciprocal, conservative depth bound).  Evaluated by the raster kernel for every list entry.
__device__ __forceinline__ void rasterConstants(const ScreenBounds& b, const FrameParams& fp, TriRecord& r)
{
	r.area = fabsf(diff2(r.x0 - r.x1, r.y0 - r.y2, r.y0 - r.y1, r.x0 - r.x2));
	// The reference's inflated 16-bit bbox (CUDAROPs.cu:180-200) decides the BINS, while the raster kernel
	// only needs a superset of the pixels whose three edge functions can be >= 0.  A covered pixel has
	// non-negative barycentrics, so it lies inside the float bounds of the triangle up to the rounding
	// error of the edge functions: a pixel delta outside the bounds has a true edge value of at most
	// -area*delta/(2*extent), against an fp32 evaluation error below 2^-20 * (delta + ext) * ext.  For
	// area > 2^-17 * D^2 (D = extents + 2) that limits delta to 2^-18 * D^2 * extent / area, which is the eps
	// below (plus the rounding of the coordinates themselves).  Needle and degenerate triangles (smaller area:
	// every pixel near the supporting LINE can evaluate to three non-negative edge functions, at any distance)
	// and non-finite ones get the whole screen, i.e. the reference's "every pixel of the bin"
	// (CUDARenderer.cu:1116-1127).
	const float extx = b.maxx - b.minx, exty = b.maxy - b.miny;
	const float D = extx + exty + 2.f;
	const float maxAbs = fmaxf(fmaxf(fabsf(b.minx), fabsf(b.maxx)), fmaxf(fabsf(b.miny), fabsf(b.maxy)));
	// hoisted reciprocal first: it also serves the margin (times 1 + 2^-20, so the product is an upper bound of
	// D*D/area; a zero / non-finite area makes eps non-finite and selects the whole-screen box below)
	r.rcpArea = refinedReciprocal(r.area);
	const float k = 3.814697265625e-6f * D * D * (r.rcpArea * 1.00000095367431640625f);
	const float epsx = 0.015625f + maxAbs * 1.52587890625e-5f + k * extx;
	const float epsy = 0.015625f + maxAbs * 1.52587890625e-5f + k * exty;
	int cx0 = 0, cx1 = (int)fp.width, cy0 = 0, cy1 = (int)fp.height;
	const bool wellConditioned = epsx < 1.0e6f && epsy < 1.0e6f && maxAbs < 1.0e9f && r.area > 7.62939453125e-6f * D * D;   // false for NaN / inf
	if (wellConditioned)
	{
		cx0 = clampi(__float2int_ru(b.minx - epsx), 0, (int)fp.width);
		cx1 = clampi(__float2int_rd(b.maxx + epsx) + 1, 0, (int)fp.width);
		cy0 = clampi(__float2int_ru(b.miny - epsy), 0, (int)fp.height);
		cy1 = clampi(__float2int_rd(b.maxy + epsy) + 1, 0, (int)fp.height);
	}
	r.cbX = (uint32_t)cx0 | ((uint32_t)cx1 << 16);
	r.cbY = (uint32_t)cy0 | ((uint32_t)cy1 << 16);
	r.divLo = divisionThreshold(r.area);
	// ... and a conservative lower bound of the integer depth.  The computed weights of a produced fragment are
	// >= 0 and, for a well-conditioned triangle, sum to 1 within k (= 2^-18 D^2 / area < 1/2, the relative error
	// of the edge functions and of the area), so z >= min(z_i) * (1 - k) for min(z_i) > 0.  A needle / degenerate
	// triangle has arbitrary (even NaN: 0/0) weights, so its fragments can have any depth - including the
	// integer 0 that cvt.rzi makes of NaN - and gets no bound; so does anything non-finite or negative.
	const float zb = fminf(fminf(r.z0, r.z1), r.z2) * (1.f - k) * 0.999999f - 1e-9f;
	r.zLow = (wellConditioned && zb > 0.f && zb < 1.f) ? __float2int_rz(zb * 2147483648.f) : 0;
	r.pad = 0;
}

// (triangle_setup_kernel lives in raster_kernel.cuh: it resolves small triangles itself and shares the fragment test)

// cooperative binning of triangles with large rectangles (full-screen quads ...): one CTA per
// big triangle, threads stride over its tiles.  SCATTER=false counts (bin_big_count_kernel), true writes
// the ids (extra CTAs of bin_scatter_kernel).
template <bool SCATTER>
__device__ __forceinline__ void binBigBody(uint32_t cta, uint32_t nCtas,
	const uint32_t* __restrict__ bigList, const uint32_t* __restrict__ counters, const uint2* __restrict__ rects,
	uint32_t* __restrict__ tileCount, const uint32_t* __restrict__ tileOffset, uint32_t* __restrict__ tileCursor,
	uint32_t* __restrict__ binIds, const FrameParams& fp, const ListMirror& M = ListMirror{ nullptr, 0u, 0u })
{
	const uint32_t nBig = counters[1];
	for (uint32_t b = cta; b < nBig; b += nCtas)
	{
		const uint32_t prim = bigList[b];
		const uint2 rc = rects[prim];
		const int bx0 = rc.x & 0xFFFF, bx1 = rc.x >> 16, by0 = rc.y & 0xFFFF, by1 = rc.y >> 16;
		const int w = bx1 - bx0 + 1, n = w * (by1 - by0 + 1);
		for (int i = threadIdx.x; i < n; i += blockDim.x)
		{
			const uint32_t tile = (uint32_t)(bx0 + i % w) + (uint32_t)(by0 + i / w) * fp.binsX;
			if (!tileOwned(fp, tile)) continue;
			if (SCATTER)
			{
				const uint32_t pos = tileOffset[tile] + atomicAdd(&tileCursor[tile], 1u);
				binIds[pos] = prim;
				mirrorStore(M, tile, pos, prim);
			}
			else
				atomicAdd(&tileCount[tile], 1u);
		}
	}
}

// counting pass of the big triangles (their scatter pass rides in bin_scatter_kernel's grid)
__global__ void __launch_bounds__(256) bin_big_count_kernel(
	const uint32_t* __restrict__ bigList, const uint32_t* __restrict__ counters, const uint2* __restrict__ rects,
	uint32_t* __restrict__ tileCount, FrameParams fp)
{
	pdlPrologue();
	binBigBody<false>(blockIdx.x, gridDim.x, bigList, counters, rects, tileCount, nullptr, nullptr, nullptr, fp);
}

// Work order of the raster kernel: tiles grouped into classes by list length, longest first (class c holds the
// tiles with at least kClassMin[c] entries; the last class the empty ones), so that its persistent CTAs take the
// long lists first and the tail of the kernel is made of short ones (longest-processing-time-first scheduling).
constexpr uint32_t kTileClasses = 6;
constexpr uint32_t kInfoClassBase = 8;       // frameInfo[8 .. 13]: tiles per class; frameInfo[14]: the raster kernel's work counter
constexpr uint32_t kInfoWorkCounter = kInfoClassBase + kTileClasses;
constexpr uint32_t kInfoDirectPairs = 5;     // frameInfo[5]: (triangle, tile) pairs of the triangles the setup kernel resolved itself (never binned)
#ifndef B2R_EMPTY_BATCH
#define B2R_EMPTY_BATCH 16
#endif
constexpr uint32_t kEmptyBatch = B2R_EMPTY_BATCH;      // empty 16x16 blocks cleared per work item (a multiple of 8: whole blocks per warp)
static_assert(kEmptyBatch % 8u == 0u && kEmptyBatch >= 8u, "a warp of the raster kernel clears whole blocks");
__device__ __forceinline__ uint32_t tileClass(uint32_t n)
{
	return n >= 1024u ? 0u : n >= 512u ? 1u : n >= 256u ? 2u : n >= 96u ? 3u : n >= 1u ? 4u : 5u;
}

// exclusive scan of the per-tile counts: a chained scan over a handful of co-resident CTAs (2048 tiles
// each; a 4K frame is 16 CTAs).  Every CTA publishes the sum of its slice, adds up the published sums of
// its predecessors (they are all resident, so spinning is safe) and scans its slice locally.  Also clears
// the cursors, publishes the pair total / overflow flag, builds the work list of the long-list sorter
// (tiles whose list is too long for the raster kernel's in-CTA sort) and the raster kernel's work order
// (tileOrder = kTileClasses lists of nTiles {tile, count, offset, 0} entries; null: not wanted - the sharded front
// end builds it from the merged heads of all ranks, shard_merge_heads_kernel).
constexpr uint32_t kScanThreads = 256, kScanPerThread = 8, kScanPerCta = kScanThreads * kScanPerThread;
__global__ void __launch_bounds__(kScanThreads) tile_scan_kernel(
	const uint32_t* __restrict__ tileCount, uint32_t* __restrict__ tileOffset, uint32_t* __restrict__ tileCursor,
	uint32_t nTiles, unsigned long long capacity, uint32_t longThreshold,
	uint32_t* __restrict__ longTiles, uint4* __restrict__ tileOrder, const uint8_t* __restrict__ dirty /*direct path: tiles whose key-buffer blocks hold fragments, or null*/,
	uint32_t ownerRank, uint32_t ownerCount, uint32_t* __restrict__ counters,
	const __grid_constant__ ShardInbox inbox, uint32_t inboxRanks /*sharded front end: the heads (count, offset) of this rank's lists go to the inbox of each tile's owner; 0: no*/,
	unsigned long long* __restrict__ frameInfo /*[0]=pairs,[1]=overflow,[2]=unsorted tiles,[3]=long tiles,[6]=lists sorted by the bitonic fallback,[8..13]=tiles per class*/)
{
	pdlPrologue();
	// published slice sums: bit 63 = ready (frame_clear_kernel zeroes them)
	volatile unsigned long long* partial = reinterpret_cast<volatile unsigned long long*>(counters + kScanBase);
	if (blockIdx.x == 0 && threadIdx.x < 32)
	{
		uint32_t v = counters[kVisBase + threadIdx.x * 32] + counters[kVisBase + (threadIdx.x + 32) * 32];
		// (triangle, tile) pairs of the triangles the setup kernel resolved itself (64-bit partial sums next to the visible counts)
		const unsigned long long* dp = reinterpret_cast<const unsigned long long*>(counters + kVisBase);
		unsigned long long d2 = dp[threadIdx.x * 16 + 1] + dp[(threadIdx.x + 32) * 16 + 1];
		for (int d = 16; d > 0; d >>= 1) { v += __shfl_down_sync(0xFFFFFFFFu, v, d); d2 += __shfl_down_sync(0xFFFFFFFFu, d2, d); }
		if (threadIdx.x == 0) { counters[0] = v; frameInfo[kInfoDirectPairs] = d2; }
	}
	__shared__ unsigned long long sWarp[kScanThreads / 32];
	__shared__ unsigned long long sPrefix;
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t first = blockIdx.x * kScanPerCta + threadIdx.x * kScanPerThread;
	uint32_t v[kScanPerThread];
	unsigned long long sum = 0;
	#pragma unroll
	for (uint32_t q = 0; q < kScanPerThread / 4; ++q)
	{
		uint4 t = make_uint4(0, 0, 0, 0);
		const uint32_t i = first + q * 4;
		if (i + 3 < nTiles) t = *reinterpret_cast<const uint4*>(tileCount + i);
		else
		{
			if (i < nTiles) t.x = tileCount[i];
			if (i + 1 < nTiles) t.y = tileCount[i + 1];
			if (i + 2 < nTiles) t.z = tileCount[i + 2];
		}
		v[q * 4] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
		sum += (unsigned long long)t.x + t.y + t.z + t.w;
	}
	unsigned long long incl = sum;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
	{
		const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
		if (lane >= (uint32_t)d) incl += t;
	}
	if (lane == 31) sWarp[warp] = incl;
	__syncthreads();
	if (warp == 0)
	{
		// block total -> publish; then the sum of the predecessors' totals
		unsigned long long w = lane < kScanThreads / 32 ? sWarp[lane] : 0ull;
		#pragma unroll
		for (int d = 1; d < (int)(kScanThreads / 32); d <<= 1)
		{
			const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, w, d);
			if (lane >= (uint32_t)d) w += t;
		}
		if (lane < kScanThreads / 32) sWarp[lane] = w;
		const unsigned long long blockTotal = __shfl_sync(0xFFFFFFFFu, w, kScanThreads / 32 - 1);
		if (lane == 0)
		{
			__threadfence();
			partial[blockIdx.x] = blockTotal | (1ull << 63);
		}
		unsigned long long before = 0;
		for (uint32_t j = lane; j < blockIdx.x; j += 32)
		{
			unsigned long long p;
			while (!((p = partial[j]) >> 63)) { }
			before += p & ~(1ull << 63);
		}
		for (int d = 16; d > 0; d >>= 1) before += __shfl_down_sync(0xFFFFFFFFu, before, d);
		if (lane == 0)
		{
			sPrefix = before;
			if (blockIdx.x == gridDim.x - 1)
			{
				const unsigned long long pairs = before + blockTotal;
				frameInfo[0] = pairs;
				frameInfo[1] = (pairs > capacity) ? 1ull : 0ull;
				tileOffset[nTiles] = (pairs > capacity) ? 0xFFFFFFFFu : 0u;
				for (uint32_t p = 0; p < inboxRanks; ++p) inbox.tileOffset[p][nTiles] = (pairs > capacity) ? 0xFFFFFFFFu : 0u;      // overflow marker, to every rank
			}
		}
	}
	__syncthreads();
	unsigned long long run = sPrefix + (warp ? sWarp[warp - 1] : 0ull) + incl - sum;
	uint32_t off[kScanPerThread];
	#pragma unroll
	for (uint32_t q = 0; q < kScanPerThread / 4; ++q)
	{
		const uint32_t i = first + q * 4;
		uint4 o;
		o.x = (uint32_t)run; run += v[q * 4];
		o.y = (uint32_t)run; run += v[q * 4 + 1];
		o.z = (uint32_t)run; run += v[q * 4 + 2];
		o.w = (uint32_t)run; run += v[q * 4 + 3];
		off[q * 4] = o.x; off[q * 4 + 1] = o.y; off[q * 4 + 2] = o.z; off[q * 4 + 3] = o.w;
		if (i + 3 < nTiles)
		{
			*reinterpret_cast<uint4*>(tileOffset + i) = o;
			*reinterpret_cast<uint4*>(tileCursor + i) = make_uint4(0, 0, 0, 0);
		}
		else
		{
			if (i < nTiles) { tileOffset[i] = o.x; tileCursor[i] = 0; }
			if (i + 1 < nTiles) { tileOffset[i + 1] = o.y; tileCursor[i + 1] = 0; }
			if (i + 2 < nTiles) { tileOffset[i + 2] = o.z; tileCursor[i + 2] = 0; }
		}
		#pragma unroll
		for (uint32_t e = 0; e < 4; ++e)
			if (i + e < nTiles && v[q * 4 + e] > longThreshold) longTiles[atomicAdd(&frameInfo[3], 1ull)] = i + e;
	}
	// Sharded front end, exchange step 1: the heads of this rank's lists go to the tiles' owners right away (posted
	// NVLink stores), so the transfer hides behind scatter + push of the lists themselves
	if (inboxRanks)
	{
		#pragma unroll
		for (uint32_t e = 0; e < kScanPerThread; ++e)
		{
			const uint32_t t = first + e;
			if (t < nTiles)
			{
				const uint32_t owner = t % inboxRanks;
				inbox.tileCount[owner][t] = v[e];
				inbox.tileOffset[owner][t] = off[e];
			}
		}
	}
	// the raster kernel's work order: one atomic per warp and class; inside a warp the tiles stay in ascending order.
	// Sort-first with a replicated front end: only the tiles this rank owns (the others hold no entries here and
	// belong to other ranks' surfaces).
	if (tileOrder)
	{
		uint32_t cls[kScanPerThread];
		// direct path: a tile whose key-buffer blocks hold fragments is a work item even with an empty list
		// (the map is padded to a multiple of 16 bytes; first is a multiple of kScanPerThread = 8)
		unsigned long long dirtyBytes = 0ull;
		if (dirty && first < nTiles) dirtyBytes = *reinterpret_cast<const unsigned long long*>(dirty + first);
		#pragma unroll
		for (uint32_t e = 0; e < kScanPerThread; ++e)
		{
			const uint32_t t = first + e;
			const bool mine = t < nTiles && (ownerCount <= 1u || t % ownerCount == ownerRank);
			const bool d = ((dirtyBytes >> (8u * e)) & 0xFFull) != 0ull;
			cls[e] = mine ? ((v[e] == 0u && d) ? kTileClasses - 2u : tileClass(v[e])) : kTileClasses;
		}
		#pragma unroll
		for (uint32_t c = 0; c < kTileClasses; ++c)
		{
			uint32_t mine = 0;
			#pragma unroll
			for (uint32_t e = 0; e < kScanPerThread; ++e) mine += (cls[e] == c) ? 1u : 0u;
			if (!__any_sync(0xFFFFFFFFu, mine != 0u)) continue;
			uint32_t inclS = mine;
			#pragma unroll
			for (int d = 1; d < 32; d <<= 1)
			{
				const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inclS, d);
				if (lane >= (uint32_t)d) inclS += t;
			}
			const uint32_t warpTotal = __shfl_sync(0xFFFFFFFFu, inclS, 31);
			uint32_t base = 0;
			if (lane == 0) base = (uint32_t)atomicAdd(&frameInfo[kInfoClassBase + c], (unsigned long long)warpTotal);
			base = __shfl_sync(0xFFFFFFFFu, base, 0) + inclS - mine;
			uint4* dst = tileOrder + (size_t)c * nTiles;
			#pragma unroll
			for (uint32_t e = 0; e < kScanPerThread; ++e)
				if (cls[e] == c) dst[base++] = make_uint4(first + e, v[e], off[e], (uint32_t)((dirtyBytes >> (8u * e)) & 0xFFull));      // the raster kernel's work item: tile, list length, list offset, key-buffer flag
		}
	}
}

#ifndef B2R_SCATTER_PER_THREAD
#define B2R_SCATTER_PER_THREAD 2
#endif
constexpr int kScatterPerThread = B2R_SCATTER_PER_THREAD;
// scatter of the small rectangles: one thread per primitive, re-reading the 8 B rectangle; the
// slot inside the tile's list comes from a warp-aggregated atomic on the tile cursor.
__global__ void __launch_bounds__(256) bin_scatter_kernel(
	const uint2* __restrict__ rects, const uint32_t* __restrict__ binSlots /*direct path: see triangle_setup_kernel; else null*/, uint32_t primLo, uint32_t primHi, const uint32_t* __restrict__ tileOffset,
	uint32_t* __restrict__ tileCursor, uint32_t* __restrict__ binIds,
	const uint32_t* __restrict__ bigList, const uint32_t* __restrict__ counters, uint32_t primBlocks,
	const unsigned long long* __restrict__ frameInfo, FrameParams fp,
	const __grid_constant__ ShardRegions mirror, uint32_t mirrorRanks /*sharded front end: number of ranks, else 0*/, uint32_t mirrorSelf)
{
	pdlPrologue();
	if (frameInfo[1]) return;   // capacity overflow: the host regrows the list buffer and re-runs the frame
	const ListMirror M{ &mirror, mirrorRanks, mirrorSelf };
	if (blockIdx.x >= primBlocks)
	{
		// the CTAs behind the per-primitive part scatter the cooperatively binned (big) triangles
		binBigBody<true>(blockIdx.x - primBlocks, gridDim.x - primBlocks, bigList, counters, rects, nullptr, tileOffset, tileCursor, binIds, fp, M);
		return;
	}
	// kScatterPerThread primitives per thread (a CTA covers kScatterPerThread * 256 consecutive ids): the rectangle
	// loads are issued together, so their latency is paid once; each warpBinRects call still sees one aligned
	// group of 32 consecutive ids, which is all the run sorter assumes.
	// Direct path: only the primitives triangle_setup_kernel left in the slots (one slot = the ids of one of its warps, in
	// lane order - so each call still sees ascending ids of one aligned group of 32); the grid is a fixed number of CTAs
	// that stride over the slots (their number is only known on the device, and is usually tiny).
	const uint32_t nSlots = binSlots ? __ldg(counters + kBinSlotsWord) : 0u;
	for (uint32_t vb = blockIdx.x; binSlots ? vb * (kScatterPerThread * 8u) < nSlots : vb == blockIdx.x; vb += primBlocks)
	{
		uint2 rc[kScatterPerThread];
		uint32_t prims[kScatterPerThread];
		#pragma unroll
		for (int k = 0; k < kScatterPerThread; ++k)
		{
			if (binSlots)
			{
				const uint32_t slot = (vb * kScatterPerThread + k) * 8u + (threadIdx.x >> 5);
				prims[k] = slot < nSlots ? __ldg(binSlots + (size_t)slot * 32u + (threadIdx.x & 31u)) : kNoPrim;
			}
			else
			{
				const uint32_t prim = primLo + (vb * kScatterPerThread + k) * blockDim.x + threadIdx.x;
				prims[k] = prim < primHi ? prim : kNoPrim;
			}
		}
		#pragma unroll
		for (int k = 0; k < kScatterPerThread; ++k)
			rc[k] = (prims[k] != kNoPrim) ? __ldg(rects + prims[k]) : make_uint2(0x0000FFFFu, 0x0000FFFFu);
		#pragma unroll
		for (int k = 0; k < kScatterPerThread; ++k)
		{
			const uint32_t prim = prims[k];
			uint32_t nTiles = 0;
			const int bx0 = rc[k].x & 0xFFFF, by0 = rc[k].y & 0xFFFF, bx1 = (rc[k].x >> 16) & 0x7FFF, by1 = rc[k].y >> 16;
			if (bx0 <= bx1 && by0 <= by1 && !(rc[k].x & kRectDirect))
			{
				nTiles = (uint32_t)(bx1 - bx0 + 1) * (uint32_t)(by1 - by0 + 1);
				if (nTiles > fp.bigRectTiles) nTiles = 0;   // binBigBody
			}
			warpBinRects<true>(nTiles, bx0, bx1, by0, by1, prim, tileCursor, tileOffset, binIds, fp, M);
		}
	}
}

// Cross-GPU ordering point of the sharded front end (one warp).  Every rank stores `epoch` into its slot of
// the peers' flag words over NVLink (signal), then (wait) spins until all ranks' slots in its OWN flag words
// have reached the epoch.  "Lists ready" (slots 0..7) does both inside every frame; "colours landed" (slots
// 8..15) is signalled by publish_info_kernel and waited for by rank 0 only when it reads or presents the frame.  The preceding kernels of the stream have completed,
// so their writes (bin lists; colours stored into rank 0's surface) are visible before the flag.  A peer
// that never arrives raises frameInfo[7] after ~4 s instead of hanging the GPU.
// earlyTrigger = 0 (frame ring): the kernel does not release its programmatic dependents before it exits - the CTAs of
// the raster kernel behind a barrier would otherwise become resident and sit out the wait for the slowest peer on
// SM slots that the front ends of the following frames (other ring streams) could use.
__global__ void __launch_bounds__(32) peer_barrier_kernel(const __grid_constant__ PeerFlags pf, uint32_t self, uint32_t n, uint32_t slotBase, uint32_t epoch,
	uint32_t signal, uint32_t wait, unsigned long long* timeoutFlag /*device frameInfo[7] or a pinned host word*/, uint32_t earlyTrigger)
{
	asm volatile("griddepcontrol.wait;" ::: "memory");
	if (earlyTrigger) asm volatile("griddepcontrol.launch_dependents;");
	const uint32_t lane = threadIdx.x;
	__threadfence_system();
	if (signal && lane < n)
		*reinterpret_cast<volatile uint32_t*>(pf.flags[lane] + slotBase + self) = epoch;
	if (wait && lane < n)
	{
		const volatile uint32_t* f = reinterpret_cast<const volatile uint32_t*>(pf.flags[self] + slotBase + lane);
		const unsigned long long t0 = globalTimerNs();
		while ((int)(*f - epoch) < 0)
		{
			if (globalTimerNs() - t0 > 4000000000ull) { *timeoutFlag = 1ull; break; }
			__nanosleep(64);
		}
	}
	__threadfence_system();
}

// per-tile sort of the id lists (ascending primitive id = submission order): makes the lists
// bit-reproducible regardless of the order the scatter atomics were served.
//  lists up to kStage entries are sorted by the CTA that RASTERISES the tile (raster_shade_kernel stages the ids of
//          its tile in shared memory anyway): ctaSortTileList below.  warpBinRects() writes one ascending run per
//          (warp, tile) and the runs of different warps cover disjoint id ranges, so the list is sorted by ORDERING
//          THE RUNS by their first key and moving them; where that does not hold (interleaved or too many runs) the
//          staged keys go through a bitonic network.  In the sharded front end the owner of a tile sorts the
//          concatenation of the ranks' segments, so the sorting work is divided with the tiles.
//  longer lists (tile_scan_kernel's "long" work list, up to kSortLarge entries) are radix-sorted in dynamic shared
//          memory, one CTA per SM, before the raster kernel (bin_sort_long_kernel)
//  lists beyond that stay in scatter order and are counted in frameInfo[2] (the image does not
//  depend on list order: the depth resolve is an order-independent min over (z, primitive id)).
//  B2R_OPT_SORT_BINS = 0 leaves all lists in scatter order; b2r_read_bins then canonicalises them when they are
//  read (bin_sort_tiles_kernel: the same CTA-level sort, one CTA per tile).
constexpr int kStage = 3072;         // <= : staged + sorted by the tile's raster CTA (12 KB of its shared memory)
constexpr int kSortLarge = 40960;    // <= : one CTA per SM, dynamic shared memory
constexpr size_t kSortLongSmem = (size_t)(kSortLarge + 32 * 256 + 2 * 32 + 2) * sizeof(uint32_t);
constexpr int kSortThreads = 256, kSortPerThread = kStage / kSortThreads;     // 12 consecutive keys per thread
static_assert(kSortPerThread % 4 == 0 && kSortPerThread <= 32, "ctaSortTileList reads its keys with 16-byte loads and keeps the run flags in one word");
constexpr int kCtaRuns = 512;        // more runs than this: bitonic fallback

// exclusive scan of one value per thread over a 256-thread CTA; *total receives the block sum
__device__ __forceinline__ uint32_t blockExclusiveScan256(uint32_t v, uint32_t* sWarp /*[9]*/, uint32_t* total)
{
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t incl = v;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= (uint32_t)d) incl += t; }
	__syncthreads();                       // sWarp may still be read from a previous call
	if (lane == 31) sWarp[warp] = incl;
	__syncthreads();
	if (threadIdx.x == 0)
	{
		uint32_t run = 0;
		for (int w = 0; w < 8; ++w) { const uint32_t c = sWarp[w]; sWarp[w] = run; run += c; }
		sWarp[8] = run;
	}
	__syncthreads();
	*total = sWarp[8];
	return sWarp[warp] + incl - v;
}

// Bitonic network over the n <= kStage keys staged in K; the sorted keys are stored through put(position, key).  All 256
// threads call it.  The network is the all-ascending ("flip") formulation - every compare-exchange moves the smaller key
// to the lower index - so n need not be a power of two: the positions from n to the next power of two hold a virtual
// +infinity and exchanges with them are skipped.  Short lists are sorted by one warp alone (no CTA barriers).
template <typename Put>
__device__ __forceinline__ void ctaBitonicSortList(uint32_t* K, uint32_t n, Put put)
{
	const uint32_t tid = threadIdx.x;
	uint32_t P = 2;
	while (P < n) P <<= 1;
	const bool oneWarp = P <= 256u;
	const uint32_t workers = oneWarp ? 32u : (uint32_t)kSortThreads;
	if (!oneWarp || tid < 32u)
		for (uint32_t k = 2; k <= P; k <<= 1)
		{
			// flip: element off of the lower half of a k-block against its mirror image in the upper half
			for (uint32_t idx = tid; idx < (P >> 1); idx += workers)
			{
				const uint32_t half = k >> 1, blockBase = (idx / half) * k, off = idx & (half - 1u);
				const uint32_t i = blockBase + off, p = blockBase + k - 1u - off;
				if (p < n)
				{
					const uint32_t x = K[i], y = K[p];
					if (x > y) { K[i] = y; K[p] = x; }
				}
			}
			if (oneWarp) __syncwarp(); else __syncthreads();
			for (uint32_t j = k >> 2; j > 0; j >>= 1)
			{
				for (uint32_t idx = tid; idx < (P >> 1); idx += workers)
				{
					const uint32_t i = ((idx & ~(j - 1u)) << 1) | (idx & (j - 1u)), p = i + j;
					if (p < n)
					{
						const uint32_t x = K[i], y = K[p];
						if (x > y) { K[i] = y; K[p] = x; }
					}
				}
				if (oneWarp) __syncwarp(); else __syncthreads();
			}
		}
	if (oneWarp) __syncthreads();
	for (uint32_t i = tid; i < n; i += kSortThreads) put(i, K[i]);
}

// The same run ordering by ONE warp, without CTA barriers: the raster kernel's warp 0 sorts the staged list of its
// tile while the other warps already sweep it (the sweep does not depend on the order).  Returns 0 = in order
// already, 1 = runs moved, 2 = not applicable (too many or interleaved runs: the caller falls back to the bitonic
// network once the sweep is over).  K is only read.
constexpr int kWarpRuns = 192;
struct WarpSortScratch
{
	uint32_t head[kWarpRuns];
	uint16_t start[kWarpRuns + 2], order[kWarpRuns], out[kWarpRuns];
};
template <typename Put>
__device__ __forceinline__ int warpSortTileList(const uint32_t* K, uint32_t n, WarpSortScratch& S, uint32_t lane, Put put)
{
	const unsigned ltMask = (1u << lane) - 1u;
	// 1. the runs.  A run = the ids one scatter warp wrote: it ends where the next key is smaller or belongs to
	//    another 32-aligned id group (warpBinRects: one ascending run per warp and tile)
	uint32_t runs = 0, carry = 0;
	bool descent = false;
	for (uint32_t b = 0; b < n; b += 32)
	{
		const uint32_t i = b + lane;
		const uint32_t k = (i < n) ? K[i] : 0xFFFFFFFFu;
		uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, k, 1);
		if (lane == 0) prev = carry;
		const bool start = (i < n) && (i == 0 || prev > k || (prev >> 5) != (k >> 5));
		descent |= (i < n) && i > 0 && prev > k;
		const unsigned m = __ballot_sync(0xFFFFFFFFu, start);
		if (start)
		{
			const uint32_t r = runs + (uint32_t)__popc(m & ltMask);
			if (r < (uint32_t)kWarpRuns) { S.start[r] = (uint16_t)i; S.head[r] = k; }
		}
		runs += (uint32_t)__popc(m);
		carry = __shfl_sync(0xFFFFFFFFu, k, 31);
	}
	if (!__any_sync(0xFFFFFFFFu, descent)) return 0;
	if (runs > (uint32_t)kWarpRuns) return 2;
	if (lane == 0) S.start[runs] = (uint16_t)n;
	__syncwarp();
	// 2. order the runs by their first key (keys are distinct primitive ids)
	for (uint32_t r = lane; r < runs; r += 32)
	{
		const uint32_t head = S.head[r];
		uint32_t rank = 0;
		for (uint32_t q = 0; q < runs; ++q) rank += (S.head[q] < head) ? 1u : 0u;
		S.order[rank] = (uint16_t)r;
	}
	__syncwarp();
	// 3. output position of every run = prefix of the run lengths in that order; the runs must cover disjoint id
	//    ranges (last key of a run < first key of the next one)
	uint32_t running = 0;
	bool bad = false;
	for (uint32_t b = 0; b < runs; b += 32)
	{
		const uint32_t kk = b + lane;
		uint32_t len = 0, r = 0;
		if (kk < runs)
		{
			r = S.order[kk];
			len = (uint32_t)S.start[r + 1] - (uint32_t)S.start[r];
			if (kk + 1 < runs) bad |= K[S.start[r + 1] - 1] > S.head[S.order[kk + 1]];
		}
		uint32_t incl = len;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= (uint32_t)d) incl += t; }
		if (kk < runs) S.out[r] = (uint16_t)(running + incl - len);
		running += __shfl_sync(0xFFFFFFFFu, incl, 31);
	}
	if (__any_sync(0xFFFFFFFFu, bad)) return 2;
	__syncwarp();
	// 4. move every key to its run's output position
	uint32_t before = 0;
	for (uint32_t b = 0; b < n; b += 32)
	{
		const uint32_t i = b + lane;
		const uint32_t k = (i < n) ? K[i] : 0u, prev = (i < n && i > 0) ? K[i - 1] : 0u;
		const bool start = (i < n) && (i == 0 || prev > k || (prev >> 5) != (k >> 5));
		const unsigned m = __ballot_sync(0xFFFFFFFFu, start);
		if (i < n)
		{
			const uint32_t r = before + (uint32_t)__popc(m & (ltMask | (1u << lane))) - 1u;
			put((uint32_t)S.out[r] + (i - (uint32_t)S.start[r]), k);
		}
		before += (uint32_t)__popc(m);
	}
	return 1;
}

// scratch of ctaSortTileList (the raster kernel overlays it on its big-triangle buffer, which is idle until the sweep)
struct __align__(16) SortScratch
{
	uint32_t head[kCtaRuns];
	uint16_t start[kCtaRuns + 2], order[kCtaRuns], out[kCtaRuns];
	uint32_t warpScan[9];
};

// Sort the n <= kStage keys staged in K (shared memory, capacity kStage, 16-byte aligned) into ascending order and
// store them through put(position, key) - the tile's list in global memory (one contiguous list, or the ranks'
// segments of the sharded front end).  All 256 threads of the CTA call it; K holds some permutation of the keys
// afterwards.  Returns 0 = the list was in order already (nothing stored), 1 = runs moved, 2 = bitonic fallback.
template <typename Put>
__device__ __forceinline__ int ctaSortTileList(uint32_t* K, uint32_t n, SortScratch& S, Put put)
{
	const uint32_t tid = threadIdx.x;
	// 1. run starts among this thread's kSortPerThread consecutive keys.  A run = the ids one scatter warp wrote: it ends
	//    where the next key is smaller or belongs to another 32-aligned id group
	const uint32_t first = tid * kSortPerThread;
	uint32_t key[kSortPerThread];
	uint32_t flags = 0;
	int descent = 0;
	if (first < n)
	{
		#pragma unroll
		for (int q = 0; q < kSortPerThread / 4; ++q)
		{
			const uint4 a = *reinterpret_cast<const uint4*>(K + first + 4 * q);
			key[4 * q] = a.x; key[4 * q + 1] = a.y; key[4 * q + 2] = a.z; key[4 * q + 3] = a.w;
		}
		uint32_t prev = first ? K[first - 1] : 0u;
		#pragma unroll
		for (uint32_t k = 0; k < (uint32_t)kSortPerThread; ++k)
		{
			const uint32_t i = first + k;
			if (i < n)
			{
				if (i == 0 || prev > key[k] || (prev >> 5) != (key[k] >> 5)) flags |= 1u << k;
				descent |= (i > 0 && prev > key[k]);
			}
			prev = key[k];
		}
	}
	if (!__syncthreads_or(descent)) return 0;
	uint32_t runs = 0;
	const uint32_t runBase = blockExclusiveScan256((uint32_t)__popc(flags), S.warpScan, &runs);
	bool fallback = runs > (uint32_t)kCtaRuns;
	if (!fallback)
	{
		for (uint32_t k = 0, r = runBase; k < (uint32_t)kSortPerThread; ++k)
			if (flags & (1u << k)) { S.start[r] = (uint16_t)(first + k); S.head[r] = key[k]; ++r; }
		if (tid == 0) S.start[runs] = (uint16_t)n;
		__syncthreads();
		// 2. order the runs by their first key (keys are distinct primitive ids)
		for (uint32_t r = tid; r < runs; r += kSortThreads)
		{
			const uint32_t head = S.head[r];
			uint32_t rank = 0;
			for (uint32_t q = 0; q < runs; ++q) rank += (S.head[q] < head) ? 1u : 0u;
			S.order[rank] = (uint16_t)r;
		}
		__syncthreads();
		// 3. output position of every run = prefix of the run lengths in that order (two runs per thread); the runs
		//    must cover disjoint id ranges (last key of a run < first key of the next one)
		uint32_t len0 = 0, len1 = 0, r0 = 0, r1 = 0;
		int bad = 0;
		const uint32_t k0 = 2 * tid, k1 = 2 * tid + 1;
		if (k0 < runs)
		{
			r0 = S.order[k0]; len0 = (uint32_t)S.start[r0 + 1] - (uint32_t)S.start[r0];
			if (k0 + 1 < runs) bad |= K[S.start[r0 + 1] - 1] > S.head[S.order[k0 + 1]];
		}
		if (k1 < runs)
		{
			r1 = S.order[k1]; len1 = (uint32_t)S.start[r1 + 1] - (uint32_t)S.start[r1];
			if (k1 + 1 < runs) bad |= K[S.start[r1 + 1] - 1] > S.head[S.order[k1 + 1]];
		}
		uint32_t total = 0;
		const uint32_t pos = blockExclusiveScan256(len0 + len1, S.warpScan, &total);
		if (k0 < runs) S.out[r0] = (uint16_t)pos;
		if (k1 < runs) S.out[r1] = (uint16_t)(pos + len0);
		fallback = __syncthreads_or(bad) != 0;
	}
	if (!fallback)
	{
		// 4. move every key to its run's output position
		#pragma unroll
		for (uint32_t k = 0; k < (uint32_t)kSortPerThread; ++k)
		{
			const uint32_t i = first + k;
			if (i < n)
			{
				const uint32_t r = runBase + (uint32_t)__popc(flags & ((2u << k) - 1u)) - 1u;
				put((uint32_t)S.out[r] + (i - (uint32_t)S.start[r]), key[k]);
			}
		}
		return 1;
	}
	// interleaved or too many runs (incoherent warps, cooperatively binned big triangles)
	ctaBitonicSortList(K, n, put);
	return 2;
}

// Block-level stable LSD radix sort (8-bit digits) of one tile list.  Keys are distinct primitive
// ids; only the bits that vary inside the list (id - min id) are sorted, so a list whose ids span
// less than 65 536 takes two passes.  Stability comes from a fixed ownership order: warp w owns a
// contiguous segment, walks it 32 keys at a time, and lanes rank themselves among the lanes with
// the same digit (__match_any_sync); per-(warp, digit) counters in shared memory are then scanned
// digit-major / warp-minor.  bufA/bufB are ping-pong buffers (bufB may be the global list itself).
template <int WARPS>
__device__ __forceinline__ void radixSortTileList(uint32_t* __restrict__ list, uint32_t n, uint32_t* bufA, uint32_t* bufB,
	uint32_t* whist /*[WARPS*256]*/, uint32_t* sRed /*[2*WARPS + 2]*/)
{
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	// load + min / max + sortedness
	uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
	int unsorted = 0;
	for (uint32_t i = tid; i < n; i += WARPS * 32)
	{
		const uint32_t k = list[i];
		bufA[i] = k;
		kmin = min(kmin, k); kmax = max(kmax, k);
		if (i + 1 < n) unsorted |= (k > list[i + 1]);
	}
	if (!__syncthreads_or(unsorted)) return;
	kmin = __reduce_min_sync(0xFFFFFFFFu, kmin); kmax = __reduce_max_sync(0xFFFFFFFFu, kmax);
	if (lane == 0) { sRed[warp] = kmin; sRed[WARPS + warp] = kmax; }
	__syncthreads();
	if (tid == 0)
	{
		uint32_t a = sRed[0], b = sRed[WARPS];
		for (int w = 1; w < WARPS; ++w) { a = min(a, sRed[w]); b = max(b, sRed[WARPS + w]); }
		sRed[2 * WARPS] = a; sRed[2 * WARPS + 1] = b;
	}
	__syncthreads();
	kmin = sRed[2 * WARPS];
	const uint32_t range = sRed[2 * WARPS + 1] - kmin;
	const int passes = range ? (32 - __clz(range) + 7) / 8 : 0;
	// warp w owns [segBegin, segEnd): contiguous, a multiple of 32 long
	const uint32_t seg = ((n + WARPS - 1) / WARPS + 31u) & ~31u;
	const uint32_t segBegin = min(warp * seg, n), segEnd = min(segBegin + seg, n);
	uint32_t* src = bufA; uint32_t* dst = bufB;
	for (int p = 0; p < passes; ++p)
	{
		const int shift = 8 * p;
		for (uint32_t i = tid; i < WARPS * 256; i += WARPS * 32) whist[i] = 0;
		__syncthreads();
		// sweep 1: per-(warp, digit) counts
		for (uint32_t b = segBegin; b < segEnd; b += 32)
		{
			const uint32_t i = b + lane;
			const bool valid = i < segEnd;
			const uint32_t d = valid ? (((src[i] - kmin) >> shift) & 255u) : (256u + lane);
			const unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
			if (valid && (int)lane == __ffs(peers) - 1) whist[warp * 256 + d] += (uint32_t)__popc(peers);
			__syncwarp();
		}
		__syncthreads();
		// exclusive scan, digit-major / warp-minor: thread d < 256 owns digit d
		uint32_t total = 0;
		if (tid < 256)
		{
			for (int w = 0; w < WARPS; ++w) { const uint32_t c = whist[w * 256 + tid]; whist[w * 256 + tid] = total; total += c; }
		}
		uint32_t incl = total;
		for (int dd = 1; dd < 32; dd <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, dd); if (lane >= (uint32_t)dd) incl += t; }
		if (tid < 256 && lane == 31) sRed[warp] = incl;
		__syncthreads();
		if (tid < 256)
		{
			uint32_t base = incl - total;
			for (uint32_t w = 0; w < warp; ++w) base += sRed[w];
			for (int w = 0; w < WARPS; ++w) whist[w * 256 + tid] += base;
		}
		__syncthreads();
		// sweep 2: rank + scatter (same order as sweep 1 => stable)
		for (uint32_t b = segBegin; b < segEnd; b += 32)
		{
			const uint32_t i = b + lane;
			const bool valid = i < segEnd;
			const uint32_t k = valid ? src[i] : 0u;
			const uint32_t d = valid ? (((k - kmin) >> shift) & 255u) : (256u + lane);
			const unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
			const int leader = __ffs(peers) - 1;
			uint32_t base = 0;
			if (valid && (int)lane == leader) { base = whist[warp * 256 + d]; whist[warp * 256 + d] = base + (uint32_t)__popc(peers); }
			base = __shfl_sync(0xFFFFFFFFu, base, leader);
			if (valid) dst[base + (uint32_t)__popc(peers & ((1u << lane) - 1u))] = k;
			__syncwarp();
		}
		__syncthreads();
		uint32_t* t = src; src = dst; dst = t;
	}
	if (src != list)
		for (uint32_t i = tid; i < n; i += WARPS * 32) list[i] = src[i];
}

// canonical order on demand (b2r_read_bins after frames rendered with B2R_OPT_SORT_BINS = 0): the raster kernel's
// CTA-level sort as a kernel of its own, one CTA per tile with 2..kStage entries
__global__ void __launch_bounds__(kSortThreads) bin_sort_tiles_kernel(
	const uint32_t* __restrict__ tileCount, const uint32_t* __restrict__ tileOffset, uint32_t* __restrict__ binIds, uint32_t nTiles,
	unsigned long long* __restrict__ frameInfo)
{
	__shared__ __align__(16) uint32_t sK[kStage];
	__shared__ SortScratch sS;
	if (frameInfo[1]) return;
	for (uint32_t tile = blockIdx.x; tile < nTiles; tile += gridDim.x)
	{
		const uint32_t n = tileCount[tile];
		if (n < 2 || n > (uint32_t)kStage) continue;
		uint32_t* list = binIds + tileOffset[tile];
		__syncthreads();
		for (uint32_t i = threadIdx.x; i < n; i += kSortThreads) sK[i] = list[i];
		__syncthreads();
		const int how = ctaSortTileList(sK, n, sS, [&](uint32_t pos, uint32_t key) { list[pos] = key; });
		if (how == 2 && threadIdx.x == 0) atomicAdd(&frameInfo[6], 1ull);
	}
}

__global__ void __launch_bounds__(1024) bin_sort_long_kernel(
	const uint32_t* __restrict__ longTiles, const uint32_t* __restrict__ tileCount, const uint32_t* __restrict__ tileOffset,
	uint32_t* __restrict__ binIds, unsigned long long* __restrict__ frameInfo)
{
	pdlPrologue();
	extern __shared__ uint32_t sDyn[];      // [kSortLarge] keys + [32*256] counters + reductions
	if (frameInfo[1]) return;
	const uint32_t nLong = (uint32_t)frameInfo[3];
	for (uint32_t w = blockIdx.x; w < nLong; w += gridDim.x)
	{
		const uint32_t tile = longTiles[w];
		const uint32_t n = tileCount[tile];
		if (n > (uint32_t)kSortLarge)
		{
			if (threadIdx.x == 0) atomicAdd(&frameInfo[2], 1ull);
			continue;
		}
		uint32_t* list = binIds + tileOffset[tile];
		radixSortTileList<32>(list, n, sDyn, list, sDyn + kSortLarge, sDyn + kSortLarge + 32 * 256);
		__syncthreads();
	}
}

}
