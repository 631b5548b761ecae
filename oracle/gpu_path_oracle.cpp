// TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference's *GPU* rendering path.
//
// Not part of the product.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline leg may load this library, and only as the checker.
//
// Parity status: PINNED against the reference itself - tests/test_ref_gpu.py (-m gpu) runs
// oracle/_ref/ref_render (the unmodified reference CUDA renderer, built from
// /root/reference by oracle/build_ref.sh) on the GPU box and compares every surface this
// file produces; tests/golden/ holds vectors generated that way (tools/make_golden.py).
// The reference has no tests or golden vectors of its own.
//
// Every function cites the reference code it follows (paths relative to
// /root/reference/CustomCUDARasterizer).  Where the reference's result depends on an
// unordered race (atomicAdd compaction order, spin-lock acquisition order) this
// restatement resolves it in submission order: mesh order, then primitive index -
// i.e. the sequentially consistent execution of the same program.
#include "oracle_math.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

using namespace orc;

namespace {

// ---- scene dump (format shared with oracle/ref_driver.cu and the product's scenefile.py)
struct SceneHeader
{
	char magic[4];
	uint32_t version;
	uint32_t width, height;
	uint32_t binsX, binsY, binW, binH, queueCap;
	uint32_t cull, sample, vis;
	float camPos[3];
	float fovDeg, nearZ, farZ, pitchDeg, yawDeg;
	uint32_t nTextures, nMeshes;
	uint32_t flags;
};

struct IVertex { float p[3], uv[2], n[3], tan[3], c[3]; };          // Vertex.h:113-145 (56 B)
struct OVertex { float p[4], uv[2], n[3], tan[3], vd[3], c[3]; };   // Vertex.h:280-327 (72 B)
static_assert(sizeof(IVertex) == 56 && sizeof(OVertex) == 72, "layout");

struct Texture { uint32_t w = 0, h = 0; std::vector<uint32_t> px; };
struct MeshTextures { const Texture* diff = nullptr; const Texture* norm = nullptr; const Texture* spec = nullptr; const Texture* gloss = nullptr; };

struct Mesh
{
	uint32_t nV = 0, nI = 0, topology = 0;
	float pos[3] = { 0, 0, 0 };
	int tex[4] = { -1, -1, -1, -1 };
	std::vector<IVertex> iv;
	std::vector<uint32_t> idx;
	M4 world;
	M3 rot;
	MeshTextures textures;
	std::vector<OVertex> ov;
	std::vector<uint8_t> visible;    // per primitive
	std::vector<int16_t> bbox;       // xMin,xMax,yMin,yMax per primitive (BoundingBox.h:4-10 order)
	uint32_t primBase = 0;
	uint32_t numVisible = 0;
};

// CUDAStructs.h:47-57 minus the replicated per-mesh texture block
struct Frag
{
	uint32_t colour32 = 0;
	float z = 0, w = 0;
	V2 uv{ 0, 0 };
	V3 n{ 0, 0, 0 }, tan{ 0, 0, 0 }, vd{ 0, 0, 0 };
	int mesh = -1;
	uint32_t prim = 0xFFFFFFFFu;
};

struct RGB { float r, g, b; };

// (Camera: oracle_math.h, shared with the CPU-path restatement)

// ---- RGBRaw.h:23-32 : truncating pack, byte0 = b, byte1 = g, byte2 = r -> 0x00RRGGBB
inline uint32_t PackColour(const RGB& c)
{
	const uint32_t b0 = cvt_rzi_u32(c.b * 255.f) & 0xFFu;
	const uint32_t b1 = cvt_rzi_u32(c.g * 255.f) & 0xFFu;
	const uint32_t b2 = cvt_rzi_u32(c.r * 255.f) & 0xFFu;
	return b0 | (b1 << 8) | (b2 << 16);
}

// ---- CUDATextureSampler.cuh:80-104
inline bool TexelCoords(const Texture& t, const V2& uv, int& x, int& y)
{
	x = cvt_rzi_s32(fmaf(uv.x, (float)t.w, 0.5f));
	y = cvt_rzi_s32(fmaf(uv.y, (float)t.h, 0.5f));
	return !(x < 0 || y < 0 || (uint32_t)x >= t.w || (uint32_t)y >= t.h);
}
inline RGB SamplePoint(const Texture& t, const V2& uv)
{
	int x, y;
	if (!TexelCoords(t, uv, x, y)) return RGB{ 1.f, 0.f, 1.f };
	const uint32_t s = t.px[(size_t)x + (size_t)y * t.w];
	return RGB{ float(s & 0xFFu) / 255.f, float((s >> 8) & 0xFFu) / 255.f, float((s >> 16) & 0xFFu) / 255.f };
}
// ---- CUDATextureSampler.cuh:106-191 : 5-tap cross in linear-index space
inline RGB SampleLinear(const Texture& t, const V2& uv)
{
	int x, y;
	if (!TexelCoords(t, uv, x, y)) return RGB{ 1.f, 0.f, 1.f };
	const int texSize = int(t.w * t.h);
	const int o = x + y * int(t.w);
	int nb[4];
	nb[0] = o - 1;          if (nb[0] < 0) nb[0] = o;
	nb[1] = o + 1;          if (nb[1] >= texSize) nb[1] = o;
	nb[2] = o + int(t.w);   if (nb[2] >= texSize) nb[2] = o;
	nb[3] = o - int(t.w);   if (nb[3] < 0) nb[3] = o;
	float r = 0.f, g = 0.f, b = 0.f;
	for (int i = 0; i < 4; ++i)
	{
		const uint32_t s = t.px[nb[i]];
		r = fmaf(float(s & 0xFFu), 0.125f, r);
		g = fmaf(float((s >> 8) & 0xFFu), 0.125f, g);
		b = fmaf(float((s >> 16) & 0xFFu), 0.125f, b);
	}
	const uint32_t s = t.px[o];
	r = fmaf(float(s & 0xFFu), 0.5f, r);
	g = fmaf(float((s >> 8) & 0xFFu), 0.5f, g);
	b = fmaf(float((s >> 16) & 0xFFu), 0.5f, b);
	const float rev = 1.0f / 255.f;   // RGBColor.h:39-43
	return RGB{ r * rev, g * rev, b * rev };
}
inline RGB Sample(const Texture& t, const V2& uv, uint32_t sampleState)
{
	return sampleState == 1 ? SampleLinear(t, uv) : SamplePoint(t, uv);
}

// ---- CUDAROPs.cu:411-470 ShadePixelSafe, evaluated the way the sm_100a build evaluates it
RGB ShadePixelSafe(const MeshTextures& tx, const Frag& f, uint32_t sampleState)
{
	const RGB D = Sample(*tx.diff, f.uv, sampleState);
	if (!tx.norm)
		return D;
	const RGB Ns = Sample(*tx.norm, f.uv, sampleState);
	// binormal = -Cross(tan, n) (:431); tangentSpaceAxis = columns {tan, binormal, n} (Matrix3.h:33-38)
	const V3 C = d_Cross(f.tan, f.n);
	const V3 B{ -C.x, -C.y, -C.z };
	const float fx = fmaf(Ns.r, 2.f, -1.f), fy = fmaf(Ns.g, 2.f, -1.f), fz = fmaf(Ns.b, 2.f, -1.f);
	const V3 N{
		fmaf(fz, f.n.x, fmaf(f.tan.x, fx, B.x * fy)),
		fmaf(fz, f.n.y, fmaf(f.tan.y, fx, B.y * fy)),
		fmaf(fz, f.n.z, fmaf(f.tan.z, fx, B.z * fy)) };
	// observedArea = Dot(-N, L), L = (.577,-.577,-.577); ClampFast(oa,0,oa) is the identity (:439, :34-40)
	const float L = 0.577f;
	const float nxL = N.x * L;
	float oa = fmaf(N.z, L, fmaf(N.y, L, -nxL));
	oa = oa / float(3.14159265358979323846);
	oa = oa * 7.0f;
	const RGB diffuse{ D.r * oa, D.g * oa, D.b * oa };
	if (!(tx.spec && tx.gloss))
		return diffuse;
	const RGB S = Sample(*tx.spec, f.uv, sampleState);
	const RGB G = Sample(*tx.gloss, f.uv, sampleState);
	// Reflect(L, N) = -L + 2*Dot(L,N)*N (Vector.h:76-78)
	float dotLN = fmaf(N.z, -L, fmaf(N.y, -L, nxL));
	dotLN = dotLN + dotLN;
	const V3 R{ fmaf(N.x, dotLN, -L), fmaf(N.y, dotLN, L), fmaf(N.z, dotLN, L) };
	float angle = fmaf(R.z, f.vd.z, fmaf(R.x, f.vd.x, f.vd.y * R.y));
	angle = (angle < 0.f) ? 0.f : angle;       // ClampFast(angle, 0, 1) (:452)
	angle = (angle > 1.f) ? 1.f : angle;
	angle = powf(angle, G.r * 25.f);
	RGB c{ fmaf(angle, S.r, diffuse.r + 0.05f), fmaf(angle, S.g, diffuse.g + 0.05f), fmaf(angle, S.b, diffuse.b + 0.05f) };
	// MaxToOne (RGBColor.h:74-79) with Max(a,b) = a<b ? b : a (MathUtilities.h:107-113)
	const float gb = (c.g < c.b) ? c.b : c.g;
	const float mx = (c.r < gb) ? gb : c.r;
	if (mx > 1.f)
	{
		const float rev = 1.0f / mx;
		c.r *= rev; c.g *= rev; c.b *= rev;
	}
	return c;
}

}

// =====================================================================================
struct OrcFrame
{
	SceneHeader h{};
	std::vector<Texture> textures;
	std::vector<Mesh> meshes;
	M4 view, proj, viewProj;
	float camPos[3], camFwd[3];
	uint32_t totalTris = 0, totalVisible = 0;
	std::vector<std::vector<uint32_t>> bins;   // global primitive ids, submission order
	std::vector<int32_t> depth;
	std::vector<Frag> frag;
	std::vector<uint32_t> framebuffer;
	std::vector<float> zbuf;
	std::vector<uint32_t> winner;
	std::vector<uint8_t> edgeMask;             // 1 = some candidate fragment sat exactly on an edge, 2 = depth tie
	std::vector<uint32_t> binFlatCounts, binFlatIds;
	uint64_t fragsCovered = 0, fragsPassed = 0;
	std::string error;
};

namespace {

bool ParseScene(OrcFrame& F, const uint8_t* blob, size_t size)
{
	size_t off = 0;
	auto take = [&](void* dst, size_t n) { if (off + n > size) return false; memcpy(dst, blob + off, n); off += n; return true; };
	if (!take(&F.h, sizeof(SceneHeader)) || memcmp(F.h.magic, "B2RS", 4) != 0 || F.h.version != 1) { F.error = "bad header"; return false; }
	F.textures.resize(F.h.nTextures);
	for (auto& t : F.textures)
	{
		uint32_t wh[2];
		if (!take(wh, 8)) { F.error = "bad texture"; return false; }
		t.w = wh[0]; t.h = wh[1];
		t.px.resize((size_t)t.w * t.h);
		if (!take(t.px.data(), t.px.size() * 4)) { F.error = "bad texture data"; return false; }
	}
	F.meshes.resize(F.h.nMeshes);
	for (auto& m : F.meshes)
	{
		uint32_t mh[3];
		if (!take(mh, 12) || !take(m.pos, 12) || !take(m.tex, 16)) { F.error = "bad mesh"; return false; }
		m.nV = mh[0]; m.nI = mh[1]; m.topology = mh[2];
		m.iv.resize(m.nV); m.idx.resize(m.nI);
		if (!take(m.iv.data(), sizeof(IVertex) * m.nV) || !take(m.idx.data(), 4 * (size_t)m.nI)) { F.error = "bad mesh data"; return false; }
		if (m.topology != 0) { F.error = "TriangleStrip is broken on the reference GPU path (CUDARenderer.cu:733-753); only TriangleList is restated"; return false; }
	}
	return true;
}

const Texture* GetTexture(const OrcFrame& F, int id)   // CUDATextureManager.cpp:21-27
{
	if (id >= 0 && (uint32_t)id < F.textures.size()) return &F.textures[id];
	return nullptr;
}

// CUDAROPs.cu:99-118 GetNDCVertex
void VertexStage(const OrcFrame& F, Mesh& m)
{
	const M4 wvp = h_Mul(F.viewProj, m.world);   // CUDARenderer.cu:242 (host)
	const V3 cam{ F.camPos[0], F.camPos[1], F.camPos[2] };
	m.ov.resize(m.nV);
	for (uint32_t i = 0; i < m.nV; ++i)
	{
		const IVertex& in = m.iv[i];
		OVertex& o = m.ov[i];
		const V3 p{ in.p[0], in.p[1], in.p[2] };
		const V4 clip = d_MulPoint(wvp, p);
		o.p[0] = clip.x / clip.w; o.p[1] = clip.y / clip.w; o.p[2] = clip.z / clip.w; o.p[3] = clip.w;
		const V3 n = d_Mul(m.rot, V3{ in.n[0], in.n[1], in.n[2] });
		const V3 t = d_Mul(m.rot, V3{ in.tan[0], in.tan[1], in.tan[2] });
		const V4 wp = d_MulPoint(m.world, p);
		V3 vd{ wp.x - cam.x, wp.y - cam.y, wp.z - cam.z };
		d_Normalize(vd);
		o.uv[0] = in.uv[0]; o.uv[1] = in.uv[1];
		o.n[0] = n.x; o.n[1] = n.y; o.n[2] = n.z;
		o.tan[0] = t.x; o.tan[1] = t.y; o.tan[2] = t.z;
		o.vd[0] = vd.x; o.vd[1] = vd.y; o.vd[2] = vd.z;
		o.c[0] = in.c[0]; o.c[1] = in.c[1]; o.c[2] = in.c[2];
	}
}

// CUDAROPs.cu:147-178
bool IsTriangleVisible(const float* a, const float* b, const float* c)
{
	const bool xOut = (a[0] < -1.f && b[0] < -1.f && c[0] < -1.f) || (a[0] > 1.f && b[0] > 1.f && c[0] > 1.f);
	const bool yOut = (a[1] < -1.f && b[1] < -1.f && c[1] < -1.f) || (a[1] > 1.f && b[1] > 1.f && c[1] > 1.f);
	const bool zOut = (a[2] < 0.f && b[2] < 0.f && c[2] < 0.f) || (a[2] > 1.f && b[2] > 1.f && c[2] > 1.f);
	return !xOut && !yOut && !zOut;
}

inline float Min3(float a, float b, float c) { float m = a; if (b < m) m = b; if (c < m) m = c; return m; }   // CUDAROPs.cu:12-21
inline float Max3(float a, float b, float c) { float m = a; if (b > m) m = b; if (c > m) m = c; return m; }   // CUDAROPs.cu:23-32
inline int16_t ClampS16(int16_t v, int16_t lo, int16_t hi) { const int16_t c = v < lo ? lo : v; return c > hi ? hi : c; }   // CUDAROPs.cu:34-40

// CUDAROPs.cu:88-97 (add, *0.5, *W - never fused in the reference build)
inline void ToScreen(const float* ndc, float W, float H, float& sx, float& sy)
{
	sx = ((ndc[0] + 1.f) * 0.5f) * W;
	sy = ((1.f - ndc[1]) * 0.5f) * H;
}

// CUDARenderer.cu:702-796 TA_Kernel
void TriangleAssembly(const OrcFrame& F, Mesh& m)
{
	const uint32_t nT = m.nI / 3;
	m.visible.assign(nT, 0);
	m.bbox.assign((size_t)nT * 4, 0);
	m.numVisible = 0;
	const V3 fwd{ F.camFwd[0], F.camFwd[1], F.camFwd[2] };
	const float W = (float)F.h.width, H = (float)F.h.height;
	for (uint32_t t = 0; t < nT; ++t)
	{
		const float* p0 = m.ov[m.idx[3 * t + 0]].p;
		const float* p1 = m.ov[m.idx[3 * t + 1]].p;
		const float* p2 = m.ov[m.idx[3 * t + 2]].p;
		if (F.h.cull == 0 || F.h.cull == 1)
		{
			const V3 a{ p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2] };
			const V3 b{ p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2] };
			V3 N = d_Cross(a, b);
			d_Normalize(N);
			// Dot(camFwd, faceNormal): the two culling branches were contracted differently by
			// the compiler (SASS of TA_Kernel @0x06f0 vs @0x0aa0).
			const float d = (F.h.cull == 0)
				? fmaf(N.z, fwd.z, fmaf(N.x, fwd.x, N.y * fwd.y))
				: fmaf(N.z, fwd.z, fmaf(N.y, fwd.y, N.x * fwd.x));
			// BackFace :760-768 culls d<=0; the live FrontFace branch :769-777 uses the SAME test
			if (d <= 0.f) continue;
		}
		if (!IsTriangleVisible(p0, p1, p2)) continue;
		m.visible[t] = 1;
		++m.numVisible;
		// bbox as TB_Kernel computes it (CUDARenderer.cu:925-933, CUDAROPs.cu:180-200)
		float x0, y0, x1, y1, x2, y2;
		ToScreen(p0, W, H, x0, y0); ToScreen(p1, W, H, x1, y1); ToScreen(p2, W, H, x2, y2);
		int16_t xMin = (int16_t)(cvt_rzi_s32(Min3(x0, x1, x2)) - 1);
		int16_t yMin = (int16_t)(cvt_rzi_s32(Min3(y0, y1, y2)) - 1);
		int16_t xMax = (int16_t)(cvt_rzi_s32(Max3(x0, x1, x2)) + 1);
		int16_t yMax = (int16_t)(cvt_rzi_s32(Max3(y0, y1, y2)) + 1);
		xMin = ClampS16(xMin, 0, (int16_t)F.h.width);
		yMin = ClampS16(yMin, 0, (int16_t)F.h.height);
		xMax = ClampS16(xMax, 0, (int16_t)F.h.width);
		yMax = ClampS16(yMax, 0, (int16_t)F.h.height);
		int16_t* bb = &m.bbox[(size_t)t * 4];
		bb[0] = xMin; bb[1] = xMax; bb[2] = yMin; bb[3] = yMax;
	}
}

inline int ClampI(int v, int lo, int hi) { const int c = v < lo ? lo : v; return c > hi ? hi : c; }

// CUDARenderer.cu:935-942
void BinRect(const OrcFrame& F, const int16_t* bb, int& bx0, int& bx1, int& by0, int& by1)
{
	bx0 = ClampI(bb[0] / (int)F.h.binW, 0, (int)F.h.binsX);
	by0 = ClampI(bb[2] / (int)F.h.binH, 0, (int)F.h.binsY);
	bx1 = ClampI(bb[1] / (int)F.h.binW, 0, (int)F.h.binsX - 1);
	by1 = ClampI(bb[3] / (int)F.h.binH, 0, (int)F.h.binsY - 1);
}

// CUDAROPs.cu:329-392 RasterizePixelAtomic + :231-251 PerformDepthTestAtomic, one pixel
struct TriSetup
{
	float x0, y0, x1, y1, x2, y2;      // screen space
	V2 eA, eB, eC;                     // v1-v0, v2-v1, v0-v2
	float area;                        // |Cross(v0-v1, v0-v2)|
	float iw0, iw1, iw2;
	const OVertex *v0, *v1, *v2;
};

inline bool TopLeft(const V2& e) { return e.y > 0.f || (e.y == 0.f && e.x < 0.f); }

void RasterPixel(OrcFrame& F, int meshIdx, uint32_t prim, const TriSetup& T, int px, int py, int fillRule)
{
	const float fx = (float)px, fy = (float)py;
	// IsPixelInTriangle (:128-145) / EdgeFunction (:120-126): Cross(pixel - v, edge)
	const float w2 = d_diff2(fx - T.x0, T.eA.y, fy - T.y0, T.eA.x);
	const float w0 = d_diff2(fx - T.x1, T.eB.y, fy - T.y1, T.eB.x);
	const float w1 = d_diff2(fx - T.x2, T.eC.y, fy - T.y2, T.eC.x);
	if (!(w0 >= 0.f && w1 >= 0.f && w2 >= 0.f))
		return;
	const size_t p = (size_t)px + (size_t)py * F.h.width;
	if (w0 == 0.f || w1 == 0.f || w2 == 0.f)
	{
		F.edgeMask[p] |= 1;
		if (fillRule == 1)
		{
			// top-left rule applied to the exact-zero case only (new design)
			if ((w2 == 0.f && !TopLeft(T.eA)) || (w0 == 0.f && !TopLeft(T.eB)) || (w1 == 0.f && !TopLeft(T.eC)))
				return;
		}
	}
	const float b0 = w0 / T.area, b1 = w1 / T.area, b2 = w2 / T.area;
	const float z = d_sum3(b0, T.v0->p[2], b1, T.v1->p[2], b2, T.v2->p[2]);
	if (z < 0.f || z > 1.f)
		return;
	++F.fragsCovered;
	const float wden = d_sum3(T.iw0, b0, T.iw1, b1, T.iw2, b2);
	const float wInterp = 1.f / wden;
	auto S = [&](float a0, float a1, float a2) { return d_sum3(b0, a0 * T.iw0, b1, a1 * T.iw1, b2, a2 * T.iw2); };
	Frag f;
	f.z = z; f.w = wInterp;
	f.uv = V2{ S(T.v0->uv[0], T.v1->uv[0], T.v2->uv[0]) * wInterp, S(T.v0->uv[1], T.v1->uv[1], T.v2->uv[1]) * wInterp };
	f.n = V3{ S(T.v0->n[0], T.v1->n[0], T.v2->n[0]) * wInterp, S(T.v0->n[1], T.v1->n[1], T.v2->n[1]) * wInterp, S(T.v0->n[2], T.v1->n[2], T.v2->n[2]) * wInterp };
	f.tan = V3{ S(T.v0->tan[0], T.v1->tan[0], T.v2->tan[0]), S(T.v0->tan[1], T.v1->tan[1], T.v2->tan[1]), S(T.v0->tan[2], T.v1->tan[2], T.v2->tan[2]) };
	f.vd = V3{ S(T.v0->vd[0], T.v1->vd[0], T.v2->vd[0]), S(T.v0->vd[1], T.v1->vd[1], T.v2->vd[1]), S(T.v0->vd[2], T.v1->vd[2], T.v2->vd[2]) };
	d_Normalize(f.vd);
	f.colour32 = PackColour(RGB{ S(T.v0->c[0], T.v1->c[0], T.v2->c[0]), S(T.v0->c[1], T.v1->c[1], T.v2->c[1]), S(T.v0->c[2], T.v1->c[2], T.v2->c[2]) });
	f.mesh = meshIdx; f.prim = prim;
	// depth: int -> float -> compare -> int (CUDAROPs.cu:386-390, :242-244)
	const int32_t zi = cvt_rzi_s32(z * 2147483648.f);
	const float zf = (float)zi;
	if (zf < (float)F.depth[p])
	{
		F.depth[p] = cvt_rzi_s32(zf);
		F.frag[p] = f;
		++F.fragsPassed;
	}
	else if (zf == (float)F.depth[p] && F.frag[p].prim != prim)
		F.edgeMask[p] |= 2;
}

void RasterMesh(OrcFrame& F, int meshIdx, int fillRule)
{
	Mesh& m = F.meshes[meshIdx];
	const uint32_t nT = m.nI / 3;
	const float W = (float)F.h.width, H = (float)F.h.height;
	// per-bin coverage of RA_PerTileKernel: 16 threads x (binDim/16) pixels (CUDARenderer.cu:1881-1882, 1059-1067)
	const int covX = 16 * int(F.h.binW / 16), covY = 16 * int(F.h.binH / 16);
	for (uint32_t t = 0; t < nT; ++t)
	{
		if (!m.visible[t]) continue;
		const int16_t* bb = &m.bbox[(size_t)t * 4];
		int bx0, bx1, by0, by1;
		BinRect(F, bb, bx0, bx1, by0, by1);
		const uint32_t prim = m.primBase + t;
		for (int bx = bx0; bx <= bx1; ++bx)         // CUDARenderer.cu:949-961 (x outer, y inner)
			for (int by = by0; by <= by1; ++by)
				F.bins[(size_t)bx + (size_t)by * F.h.binsX].push_back(prim);
		if (bx0 > bx1 || by0 > by1) continue;

		TriSetup T;
		T.v0 = &m.ov[m.idx[3 * t + 0]]; T.v1 = &m.ov[m.idx[3 * t + 1]]; T.v2 = &m.ov[m.idx[3 * t + 2]];
		ToScreen(T.v0->p, W, H, T.x0, T.y0); ToScreen(T.v1->p, W, H, T.x1, T.y1); ToScreen(T.v2->p, W, H, T.x2, T.y2);
		T.eA = V2{ T.x1 - T.x0, T.y1 - T.y0 };
		T.eB = V2{ T.x2 - T.x1, T.y2 - T.y1 };
		T.eC = V2{ T.x0 - T.x2, T.y0 - T.y2 };
		T.area = fabsf(d_Cross2(V2{ T.x0 - T.x1, T.y0 - T.y1 }, V2{ T.x0 - T.x2, T.y0 - T.y2 }));
		T.iw0 = 1.f / T.v0->p[3]; T.iw1 = 1.f / T.v1->p[3]; T.iw2 = 1.f / T.v2->p[3];

		// Every pixel of every overlapped bin is tested in the reference (no bbox intersection,
		// CUDARenderer.cu:1116-1127) and coverage is decided by the edge test alone.  The scan below is the same
		// full-bin scan, narrowed ONLY for well-conditioned triangles: a pixel delta outside the float bounds has
		// a true edge value <= -area*delta/(2*extent) while the fp32 evaluation error is below 2^-20*(delta+ext)*ext,
		// so with area > 2^-10*D^2 (D = extents + 2) nothing farther than 1 + ext/256 pixels outside the bounds can
		// pass.  Degenerate and needle triangles (zero or tiny area: every pixel on the supporting line has all
		// three edge functions == 0 or rounding-sized) get the full scan, like in the reference.
		int X0 = bx0 * (int)F.h.binW, X1 = bx1 * (int)F.h.binW + covX;
		int Y0 = by0 * (int)F.h.binH, Y1 = by1 * (int)F.h.binH + covY;
		X1 = std::min(X1, (int)F.h.width); Y1 = std::min(Y1, (int)F.h.height);
		const float mnx = Min3(T.x0, T.x1, T.x2), mxx = Max3(T.x0, T.x1, T.x2);
		const float mny = Min3(T.y0, T.y1, T.y2), mxy = Max3(T.y0, T.y1, T.y2);
		if (std::isfinite(mnx) && std::isfinite(mxx) && std::isfinite(mny) && std::isfinite(mxy) && std::isfinite(T.area) &&
			std::fabs(mnx) < 1.0e7f && std::fabs(mxx) < 1.0e7f && std::fabs(mny) < 1.0e7f && std::fabs(mxy) < 1.0e7f)
		{
			const double ext = std::max((double)mxx - mnx, (double)mxy - mny), D = ((double)mxx - mnx) + ((double)mxy - mny) + 2.0;
			if ((double)T.area > D * D / 1024.0)
			{
				const int margin = 1 + (int)std::ceil(ext / 256.0);
				X0 = std::max(X0, (int)std::floor((double)mnx) - margin);
				X1 = std::min(X1, (int)std::floor((double)mxx) + 1 + margin);
				Y0 = std::max(Y0, (int)std::floor((double)mny) - margin);
				Y1 = std::min(Y1, (int)std::floor((double)mxy) + 1 + margin);
			}
		}
		for (int y = Y0; y < Y1; ++y)
		{
			const int inBinY = y - (y / (int)F.h.binH) * (int)F.h.binH;
			if (inBinY >= covY) continue;
			for (int x = X0; x < X1; ++x)
			{
				const int inBinX = x - (x / (int)F.h.binW) * (int)F.h.binW;
				if (inBinX >= covX) continue;
				RasterPixel(F, meshIdx, prim, T, x, y, fillRule);
			}
		}
	}
}

// CUDARenderer.cu:833-897 PS_Kernel / PS_VisualiseDepthColourKernel / PS_VisualiseNormalKernel
uint32_t ShadePixel(const OrcFrame& F, const Frag& f)
{
	switch (F.h.vis)
	{
	case 1:
	{
		const float r = (f.z - 0.985f) / (1.f - 0.985f);
		return (cvt_rzi_u32(r * 255.f) & 0xFFu) << 16;
	}
	case 2:
		return PackColour(RGB{ f.n.x, f.n.y, f.n.z });
	default:
		if (f.mesh >= 0 && F.meshes[f.mesh].textures.diff)
			return PackColour(ShadePixelSafe(F.meshes[f.mesh].textures, f, F.h.sample));
		return f.colour32;
	}
}

}

// =====================================================================================
extern "C" {

// fillRule: 0 = reference coverage (all three edge functions >= 0), 1 = top-left on exact zeros
// modelHoles: 1 = reproduce the integer-division launch grids of Clear / PixelShader
//             (CUDARenderer.cu:1642,1656,1950), 0 = every pixel cleared and shaded
OrcFrame* orc_render(const uint8_t* blob, uint64_t size, int fillRule, int modelHoles)
{
	OrcFrame* Fp = new OrcFrame;
	OrcFrame& F = *Fp;
	if (!ParseScene(F, blob, (size_t)size))
		return Fp;
	const SceneHeader& h = F.h;
	// faithful mode refuses grids the reference would write out of bounds with; with modelHoles = 0
	// (the new design's ceil grids) bins are clipped to the screen instead
	if (modelHoles && ((uint64_t)h.binsX * h.binW > h.width || (uint64_t)h.binsY * h.binH > h.height))
	{
		F.error = "bin grid exceeds the screen: RA_PerTileKernel has no screen-bounds guard (CUDARenderer.cu:1053-1067)";
		return Fp;
	}

	// camera (main.cpp:421-425 call sequence)
	Camera cam(h.camPos, h.fovDeg);
	cam.SetAspectRatio(float(h.width), float(h.height));
	cam.SetNearFar(h.nearZ, h.farZ);
	if (h.pitchDeg != 0.f) cam.Pitch(h.pitchDeg);
	if (h.yawDeg != 0.f) cam.Yaw(h.yawDeg);
	F.view = h_Inverse(cam.lookAt);                 // Camera.h:26
	F.proj = cam.proj;
	F.viewProj = h_Mul(F.proj, F.view);             // CUDARenderer.cu:219
	for (int i = 0; i < 3; ++i) { F.camPos[i] = cam.lookAt.d[3][i]; F.camFwd[i] = cam.lookAt.d[2][i]; }

	// meshes: world = Mesh world (Mesh.cpp:15-30) re-translated by CUDAMesh (CUDAMesh.cpp:4-20)
	uint32_t primBase = 0;
	for (auto& m : F.meshes)
	{
		m.world = Identity4();
		for (int i = 0; i < 3; ++i) m.world.d[3][i] += m.pos[i];    // Mesh ctor
		for (int i = 0; i < 3; ++i) m.world.d[3][i] += m.pos[i];    // CUDAMesh ctor adds it again
		for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) m.rot.d[c][r] = m.world.d[c][r];
		// CUDARenderer.cu:550-573 texture wiring
		const Texture* d = GetTexture(F, m.tex[0]); const Texture* n = GetTexture(F, m.tex[1]);
		const Texture* s = GetTexture(F, m.tex[2]); const Texture* g = GetTexture(F, m.tex[3]);
		if (d) { m.textures.diff = d; if (n) { m.textures.norm = n; if (s) m.textures.spec = s; if (g) m.textures.gloss = g; } }
		m.primBase = primBase;
		primBase += m.nI / 3;
	}
	F.totalTris = primBase;

	const size_t nPix = (size_t)h.width * h.height;
	// CUDAWindowHelper.h:9-31 buffers start zeroed; Clear (CUDARenderer.cu:1626-1659)
	F.depth.assign(nPix, 0);
	F.frag.assign(nPix, Frag{});
	F.framebuffer.assign(nPix, 0u);
	F.edgeMask.assign(nPix, 0);
	const uint32_t clearColour = PackColour(RGB{ 0.25f, 0.25f, 0.25f });   // CUDARenderer.h:135
	const uint32_t clrW = modelHoles ? 32 * (h.width / 32) : h.width, clrH = modelHoles ? 32 * (h.height / 32) : h.height;
	for (uint32_t y = 0; y < h.height; ++y)
		for (uint32_t x = 0; x < h.width; ++x)
			if (x < clrW && y < clrH)
			{
				F.depth[(size_t)x + (size_t)y * h.width] = INT32_MAX;
				F.frag[(size_t)x + (size_t)y * h.width].colour32 = clearColour;
			}

	F.bins.assign((size_t)h.binsX * h.binsY, {});
	for (size_t mi = 0; mi < F.meshes.size(); ++mi)
	{
		VertexStage(F, F.meshes[mi]);
		TriangleAssembly(F, F.meshes[mi]);
		F.totalVisible += F.meshes[mi].numVisible;
		RasterMesh(F, (int)mi, fillRule);
	}

	const uint32_t psW = modelHoles ? 16 * (h.width / 16) : h.width, psH = modelHoles ? 16 * (h.height / 16) : h.height;
	F.zbuf.resize(nPix); F.winner.resize(nPix);
	for (uint32_t y = 0; y < h.height; ++y)
		for (uint32_t x = 0; x < h.width; ++x)
		{
			const size_t p = (size_t)x + (size_t)y * h.width;
			F.zbuf[p] = F.frag[p].z;
			F.winner[p] = F.frag[p].prim;
			if (x < psW && y < psH)
				F.framebuffer[p] = ShadePixel(F, F.frag[p]);
		}

	F.binFlatCounts.resize(F.bins.size());
	for (size_t b = 0; b < F.bins.size(); ++b)
	{
		F.binFlatCounts[b] = (uint32_t)F.bins[b].size();
		F.binFlatIds.insert(F.binFlatIds.end(), F.bins[b].begin(), F.bins[b].end());
	}
	return Fp;
}

void orc_free(OrcFrame* F) { delete F; }
const char* orc_error(const OrcFrame* F) { return F->error.empty() ? nullptr : F->error.c_str(); }
uint32_t orc_width(const OrcFrame* F) { return F->h.width; }
uint32_t orc_height(const OrcFrame* F) { return F->h.height; }
uint32_t orc_num_meshes(const OrcFrame* F) { return (uint32_t)F->meshes.size(); }
uint32_t orc_num_bins(const OrcFrame* F) { return (uint32_t)F->bins.size(); }
uint32_t orc_total_tris(const OrcFrame* F) { return F->totalTris; }
uint32_t orc_total_visible(const OrcFrame* F) { return F->totalVisible; }
uint64_t orc_frags_covered(const OrcFrame* F) { return F->fragsCovered; }
uint64_t orc_frags_passed(const OrcFrame* F) { return F->fragsPassed; }
void orc_matrices(const OrcFrame* F, float* viewProj16, float* view16, float* proj16, float* camPos3, float* camFwd3)
{
	memcpy(viewProj16, F->viewProj.d, 64); memcpy(view16, F->view.d, 64); memcpy(proj16, F->proj.d, 64);
	memcpy(camPos3, F->camPos, 12); memcpy(camFwd3, F->camFwd, 12);
}
void orc_mesh_world(const OrcFrame* F, uint32_t m, float* world16) { memcpy(world16, F->meshes[m].world.d, 64); }
uint32_t orc_mesh_num_vertices(const OrcFrame* F, uint32_t m) { return F->meshes[m].nV; }
uint32_t orc_mesh_num_tris(const OrcFrame* F, uint32_t m) { return F->meshes[m].nI / 3; }
uint32_t orc_mesh_prim_base(const OrcFrame* F, uint32_t m) { return F->meshes[m].primBase; }
const float* orc_mesh_overtex(const OrcFrame* F, uint32_t m) { return reinterpret_cast<const float*>(F->meshes[m].ov.data()); }
const uint8_t* orc_mesh_visible(const OrcFrame* F, uint32_t m) { return F->meshes[m].visible.data(); }
const int16_t* orc_mesh_bbox(const OrcFrame* F, uint32_t m) { return F->meshes[m].bbox.data(); }
const uint32_t* orc_bin_counts(const OrcFrame* F) { return F->binFlatCounts.data(); }
const uint32_t* orc_bin_ids(const OrcFrame* F) { return F->binFlatIds.data(); }
uint64_t orc_bin_total(const OrcFrame* F) { return F->binFlatIds.size(); }
const uint32_t* orc_framebuffer(const OrcFrame* F) { return F->framebuffer.data(); }
const int32_t* orc_depth(const OrcFrame* F) { return F->depth.data(); }
const float* orc_zbuf(const OrcFrame* F) { return F->zbuf.data(); }
const uint32_t* orc_winner(const OrcFrame* F) { return F->winner.data(); }
const uint8_t* orc_edge_mask(const OrcFrame* F) { return F->edgeMask.data(); }
// 13 floats per pixel, same order as oracle/ref_driver.cu dumps: w, uv2, n3, tan3, vd3, colour32(bits)
void orc_frag_attrs(const OrcFrame* F, float* out13)
{
	const size_t nPix = F->frag.size();
	for (size_t p = 0; p < nPix; ++p)
	{
		const Frag& f = F->frag[p];
		float* a = out13 + p * 13;
		a[0] = f.w; a[1] = f.uv.x; a[2] = f.uv.y; a[3] = f.n.x; a[4] = f.n.y; a[5] = f.n.z;
		a[6] = f.tan.x; a[7] = f.tan.y; a[8] = f.tan.z; a[9] = f.vd.x; a[10] = f.vd.y; a[11] = f.vd.z;
		memcpy(&a[12], &f.colour32, 4);
	}
}

}
