// TEST / MEASUREMENT INFRASTRUCTURE (oracle) -- CPU restatement of the reference's software
// renderer Elite::Renderer (CustomCUDARasterizer/ERenderer.cpp), the path BASELINE.md times as
// the CPU baseline.  Not part of the product; only bench.py's cpu_baseline / --impl reference
// legs and tests/ use it.
//
// Why a restatement: the in-tree ERenderer.cpp is stale against its own headers (GetVertexAmount / GetIndexAmount /
// GetTextures / GetWorldMatrix / IsDepthColour do not exist any more; the file is excluded from the project,
// vcxproj:126-155), so it does not build with the reference's build system - and the restatement travels in this
// repository, which the reference's sources must not.
// Parity status: PINNED.  oracle/build_ref_erenderer.sh compiles the reference's ERenderer.cpp UNMODIFIED (with its own
// Camera / Light / SceneGraph / Texture sources and math headers) against a reconstruction of the two stale interfaces
// (oracle/ere_shim/) into oracle/_ref/libere_ref.so; tests/test_cpu_erenderer.py compares this restatement with it bit
// for bit (back buffer and float depth buffer) on 21 scenes incl. the bundled vehicle and BASELINE's full-size scenes,
// and against committed outputs of that build (tests/golden/erenderer_reference.npz, tools/make_golden.py erenderer).
//
// Single-threaded, immediate mode: every depth-passing fragment is shaded (ERenderer.cpp:269-430).
// Host arithmetic: plain IEEE single precision in source order (compile with -ffp-contract=off).
#include "oracle_math.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

using namespace orc;

namespace {

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
struct IVertex { float p[3], uv[2], n[3], tan[3], c[3]; };
struct OVertex { V4 p; V2 uv; V3 n, tan, vd, c; };
struct Texture { uint32_t w = 0, h = 0; std::vector<uint32_t> px; };
struct Mesh { uint32_t nV = 0, nI = 0; float pos[3]; int tex[4]; std::vector<IVertex> iv; std::vector<uint32_t> idx; M4 world; };
struct RGB { float r, g, b; };

inline RGB operator*(const RGB& a, const RGB& b) { return { a.r * b.r, a.g * b.g, a.b * b.b }; }
inline RGB operator*(const RGB& a, float f) { return { a.r * f, a.g * f, a.b * f }; }
inline RGB operator+(const RGB& a, const RGB& b) { return { a.r + b.r, a.g + b.g, a.b + b.b }; }

// Texture.h:32-45
RGB SamplePoint(const Texture& t, const V2& uv)
{
	const int x = int(uv.x * t.w + 0.5f), y = int(uv.y * t.h + 0.5f);
	if (x < 0 || y < 0 || x >= (int)t.w || y >= (int)t.h) return { 1.f, 0.f, 1.f };
	const uint32_t s = t.px[(size_t)x + (size_t)y * t.w];
	return { (s & 0xFF) / 255.f, ((s >> 8) & 0xFF) / 255.f, ((s >> 16) & 0xFF) / 255.f };
}
// Texture.h:47-124 (note the `>` bounds, :51)
RGB SampleLinear(const Texture& t, const V2& uv)
{
	const int x = int(uv.x * t.w + 0.5f), y = int(uv.y * t.h + 0.5f);
	if (x < 0 || y < 0 || x > (int)t.w || y > (int)t.h) return { 1.f, 0.f, 1.f };
	const int texSize = int(t.w * t.h);
	int o = x + y * int(t.w);
	if (o >= texSize) o = texSize - 1;      // the reference would read out of bounds here; keep the run defined
	int nb[4] = { o - 1, o + 1, o + int(t.w), o - int(t.w) };
	if (nb[0] < 0) nb[0] = o;
	if (nb[1] >= texSize) nb[1] = o;
	if (nb[2] >= texSize) nb[2] = o;
	if (nb[3] < 0) nb[3] = o;
	const float weight = 1.f / (4 * 2);
	RGB c{ 0, 0, 0 };
	for (int i = 0; i < 4; ++i)
	{
		const uint32_t s = t.px[nb[i]];
		c.r += (s & 0xFF) * weight; c.g += ((s >> 8) & 0xFF) * weight; c.b += ((s >> 16) & 0xFF) * weight;
	}
	const uint32_t s = t.px[o];
	c.r += (s & 0xFF) / 2.f; c.g += ((s >> 8) & 0xFF) / 2.f; c.b += ((s >> 16) & 0xFF) / 2.f;
	const float rev = 1.0f / 255.f;
	return { c.r * rev, c.g * rev, c.b * rev };
}
RGB Sample(const Texture& t, const V2& uv, uint32_t state) { return state == 1 ? SampleLinear(t, uv) : SamplePoint(t, uv); }

inline float Clamp01(float a) { if (a < 0.f) return 0.f; if (a > 1.f) return 1.f; return a; }

struct Frame
{
	SceneHeader h;
	std::vector<Texture> textures;
	std::vector<Mesh> meshes;
	M4 view, proj, viewProj;
	V3 camPos, camFwd;
	std::vector<uint32_t> back;
	std::vector<float> depth;
	uint64_t fragsShaded = 0, trisDrawn = 0;
};

const Texture* Tex(const Frame& F, int id) { return (id >= 0 && (size_t)id < F.textures.size()) ? &F.textures[id] : nullptr; }

// ERenderer.cpp:215-242 GetNDCVertex (recomputes the WVP product per vertex, like the reference)
OVertex NDCVertex(const Frame& F, const IVertex& v, const M4& world)
{
	const M4 wvp = h_Mul(F.viewProj, world);
	auto mulPoint = [](const M4& m, const float* p) {
		return V4{ m(0, 0) * p[0] + m(0, 1) * p[1] + m(0, 2) * p[2] + m(0, 3),
		           m(1, 0) * p[0] + m(1, 1) * p[1] + m(1, 2) * p[2] + m(1, 3),
		           m(2, 0) * p[0] + m(2, 1) * p[1] + m(2, 2) * p[2] + m(2, 3),
		           m(3, 0) * p[0] + m(3, 1) * p[1] + m(3, 2) * p[2] + m(3, 3) };
	};
	OVertex o;
	o.p = mulPoint(wvp, v.p);
	o.p.x /= o.p.w; o.p.y /= o.p.w; o.p.z /= o.p.w;
	const V4 wp = mulPoint(world, v.p);
	o.vd = V3{ wp.x - F.camPos.x, wp.y - F.camPos.y, wp.z - F.camPos.z };
	h_Normalize(o.vd);
	M3 rot; for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) rot.d[c][r] = world.d[c][r];
	o.n = h_Mul(rot, V3{ v.n[0], v.n[1], v.n[2] });
	o.tan = h_Mul(rot, V3{ v.tan[0], v.tan[1], v.tan[2] });
	o.uv = V2{ v.uv[0], v.uv[1] };
	o.c = V3{ v.c[0], v.c[1], v.c[2] };
	return o;
}

inline float Cross2(float ax, float ay, float bx, float by) { return ax * by - ay * bx; }

// ERenderer.cpp:269-430 RenderPixelsInTriangle
void RenderTriangle(Frame& F, const Mesh& m, const OVertex& v0, const OVertex& v1, const OVertex& v2)
{
	const float W = (float)F.h.width, H = (float)F.h.height;
	// NDCToScreenSpace, ERenderer.cpp:603-607
	V4 r0 = v0.p, r1 = v1.p, r2 = v2.p;
	r0.x = ((r0.x + 1) / 2) * W; r0.y = ((1 - r0.y) / 2) * H;
	r1.x = ((r1.x + 1) / 2) * W; r1.y = ((1 - r1.y) / 2) * H;
	r2.x = ((r2.x + 1) / 2) * W; r2.y = ((1 - r2.y) / 2) * H;
	// GetBoundingBox, ERenderer.cpp:609-624
	int xMin = (short)std::min({ r0.x, r1.x, r2.x }) - 1, yMin = (short)std::min({ r0.y, r1.y, r2.y }) - 1;
	int xMax = (short)std::max({ r0.x, r1.x, r2.x }) + 1, yMax = (short)std::max({ r0.y, r1.y, r2.y }) + 1;
	xMin = (short)xMin; yMin = (short)yMin; xMax = (short)xMax; yMax = (short)yMax;
	if (xMin < 0) xMin = 0;
	if (yMin < 0) yMin = 0;
	if (xMax > (int)F.h.width) xMax = (int)F.h.width;
	if (yMax > (int)F.h.height) yMax = (int)F.h.height;

	const float iw0 = 1.f / r0.w, iw1 = 1.f / r1.w, iw2 = 1.f / r2.w;
	const Texture* diff = Tex(F, m.tex[0]); const Texture* norm = Tex(F, m.tex[1]);
	const Texture* spec = Tex(F, m.tex[2]); const Texture* gloss = Tex(F, m.tex[3]);
	// IsPixelInTriangle, ERenderer.cpp:497-542
	const float eAx = r0.x - r1.x, eAy = r0.y - r1.y, eBx = r1.x - r2.x, eBy = r1.y - r2.y, eCx = r2.x - r0.x, eCy = r2.y - r0.y;
	const float totalArea = std::fabs(Cross2(eAx, eAy, eCx, eCy));
	++F.trisDrawn;
	for (int r = yMin; r < yMax; ++r)
		for (int c = xMin; c < xMax; ++c)
		{
			const float px = float(c), py = float(r);
			float w[3];
			float cr = Cross2(eAx, eAy, px - r0.x, py - r0.y);
			if (cr < 0.f) continue;
			w[2] = cr / totalArea;
			cr = Cross2(eBx, eBy, px - r1.x, py - r1.y);
			if (cr < 0.f) continue;
			w[0] = cr / totalArea;
			cr = Cross2(eCx, eCy, px - r2.x, py - r2.y);
			if (cr < 0.f) continue;
			w[1] = cr / totalArea;
			// DepthTest, ERenderer.cpp:544-554 (buffer cleared to NaN by memset(0xFF), :633-636)
			const size_t pid = (size_t)c + (size_t)r * F.h.width;
			const float z = (w[0] * r0.z) + (w[1] * r1.z) + (w[2] * r2.z);
			if (z < 0 || z > 1.f) continue;
			if (z > F.depth[pid]) continue;
			F.depth[pid] = z;
			++F.fragsShaded;

			const float wInterp = 1.f / (iw0 * w[0] + iw1 * w[1] + iw2 * w[2]);
			auto S = [&](float a0, float a1, float a2) { return w[0] * (a0 * iw0) + w[1] * (a1 * iw1) + w[2] * (a2 * iw2); };
			V2 uv{ S(v0.uv.x, v1.uv.x, v2.uv.x), S(v0.uv.y, v1.uv.y, v2.uv.y) };
			uv.x *= wInterp; uv.y *= wInterp;
			RGB fin{ 0, 0, 0 };
			if (F.h.vis != 1)
			{
				if (diff)
				{
					const RGB dc = Sample(*diff, uv, F.h.sample);
					if (norm)
					{
						const RGB nrgb = Sample(*norm, uv, F.h.sample);
						V3 n{ S(v0.n.x, v1.n.x, v2.n.x), S(v0.n.y, v1.n.y, v2.n.y), S(v0.n.z, v1.n.z, v2.n.z) };
						n = h_Scale(n, wInterp);
						const V3 t{ S(v0.tan.x, v1.tan.x, v2.tan.x), S(v0.tan.y, v1.tan.y, v2.tan.y), S(v0.tan.z, v1.tan.z, v2.tan.z) };
						const V3 b = h_Cross(t, n);
						const V3 s{ 2.f * nrgb.r - 1.f, 2.f * nrgb.g - 1.f, 2.f * nrgb.b - 1.f };
						const V3 N{ t.x * s.x + b.x * s.y + n.x * s.z, t.y * s.x + b.y * s.y + n.y * s.z, t.z * s.x + b.z * s.y + n.z * s.z };
						V3 vd{ S(v0.vd.x, v1.vd.x, v2.vd.x), S(v0.vd.y, v1.vd.y, v2.vd.y), S(v0.vd.z, v1.vd.z, v2.vd.z) };
						h_Normalize(vd);
						// one DirectionalLight{white, 2.0, (.577,-.577,-.577)} (main.cpp:178), ERenderer.cpp:356-386
						const V3 L{ 0.577f, -0.577f, -0.577f };
						const float oa = h_Dot(V3{ -N.x, -N.y, -N.z }, L);
						if (!(oa < 0.f))
						{
							const RGB bi{ 2.f, 2.f, 2.f };
							if (spec && gloss)
							{
								const float d2 = 2.f * h_Dot(L, N);
								const V3 R{ -L.x + d2 * N.x, -L.y + d2 * N.y, -L.z + d2 * N.z };
								const float angle = h_Dot(R, vd);
								const RGB ss = Sample(*spec, uv, F.h.sample);
								const RGB phong = ss * powf(angle, Sample(*gloss, uv, F.h.sample).r * 25.f);
								fin = fin + bi * (dc + phong) * oa;
							}
							else
								fin = fin + bi * dc * oa;
						}
						fin = RGB{ Clamp01(fin.r), Clamp01(fin.g), Clamp01(fin.b) };
					}
					else
						fin = dc;
				}
				else
					fin = RGB{ v0.c.x * w[0] + v1.c.x * w[1] + v2.c.x * w[2], v0.c.y * w[0] + v1.c.y * w[1] + v2.c.y * w[2], v0.c.z * w[0] + v1.c.z * w[1] + v2.c.z * w[2] };
			}
			else
			{
				fin = RGB{ Clamp01((z - 0.985f) / (1.f - 0.985f)), 0.f, 0.f };
			}
			// SDL_MapRGB on the 32-bit XRGB back buffer (ERenderer.cpp:422-425)
			F.back[pid] = ((uint32_t)(uint8_t)(int)(fin.r * 255.f) << 16) | ((uint32_t)(uint8_t)(int)(fin.g * 255.f) << 8) | (uint32_t)(uint8_t)(int)(fin.b * 255.f);
		}
}

// ERenderer.cpp:42-164
void RenderFrame(Frame& F, uint32_t meshBegin, uint32_t meshEnd, uint32_t triBegin, uint32_t triEnd)
{
	const size_t nPix = (size_t)F.h.width * F.h.height;
	memset(F.depth.data(), 0xFF, nPix * 4);          // ClearDepthBuffer: NaN floats
	memset(F.back.data(), 64, nPix * 4);             // ClearScreen
	for (uint32_t mi = meshBegin; mi < meshEnd && mi < F.meshes.size(); ++mi)
	{
		const Mesh& m = F.meshes[mi];
		std::vector<OVertex> ndc;
		ndc.reserve(m.nV);
		for (uint32_t i = 0; i < m.nV; ++i) ndc.push_back(NDCVertex(F, m.iv[i], m.world));
		const uint32_t nT = m.nI / 3;
		for (uint32_t t = triBegin; t < std::min(nT, triEnd); ++t)
		{
			const OVertex& v0 = ndc[m.idx[3 * t]]; const OVertex& v1 = ndc[m.idx[3 * t + 1]]; const OVertex& v2 = ndc[m.idx[3 * t + 2]];
			if (F.h.cull == 0 || F.h.cull == 1)
			{
				V3 N = h_Cross(V3{ v1.p.x - v0.p.x, v1.p.y - v0.p.y, v1.p.z - v0.p.z }, V3{ v2.p.x - v0.p.x, v2.p.y - v0.p.y, v2.p.z - v0.p.z });
				h_Normalize(N);
				const float d = h_Dot(F.camFwd, N);
				if (F.h.cull == 0 && d <= 0.f) continue;      // :82-91
				if (F.h.cull == 1 && d >= 0.f) continue;      // :92-102
			}
			const bool xOut = (v0.p.x < -1.f && v1.p.x < -1.f && v2.p.x < -1.f) || (v0.p.x > 1.f && v1.p.x > 1.f && v2.p.x > 1.f);
			const bool yOut = (v0.p.y < -1.f && v1.p.y < -1.f && v2.p.y < -1.f) || (v0.p.y > 1.f && v1.p.y > 1.f && v2.p.y > 1.f);
			const bool zOut = (v0.p.z < 0.f && v1.p.z < 0.f && v2.p.z < 0.f) || (v0.p.z > 1.f && v1.p.z > 1.f && v2.p.z > 1.f);
			if (xOut || yOut || zOut) continue;
			RenderTriangle(F, m, v0, v1, v2);
		}
	}
}

}

extern "C" {

struct EreScene { Frame F; std::string error; };

EreScene* ere_load(const uint8_t* blob, uint64_t size)
{
	EreScene* S = new EreScene;
	Frame& F = S->F;
	size_t off = 0;
	auto take = [&](void* dst, size_t n) { if (off + n > size) return false; memcpy(dst, blob + off, n); off += n; return true; };
	if (!take(&F.h, sizeof(SceneHeader)) || memcmp(F.h.magic, "B2RS", 4) != 0) { S->error = "bad header"; return S; }
	F.textures.resize(F.h.nTextures);
	for (auto& t : F.textures)
	{
		uint32_t wh[2]; take(wh, 8); t.w = wh[0]; t.h = wh[1]; t.px.resize((size_t)t.w * t.h); take(t.px.data(), t.px.size() * 4);
	}
	F.meshes.resize(F.h.nMeshes);
	for (auto& m : F.meshes)
	{
		uint32_t mh[3]; take(mh, 12); take(m.pos, 12); take(m.tex, 16);
		m.nV = mh[0]; m.nI = mh[1]; m.iv.resize(m.nV); m.idx.resize(m.nI);
		if (!take(m.iv.data(), sizeof(IVertex) * m.nV) || !take(m.idx.data(), 4 * (size_t)m.nI)) { S->error = "bad mesh"; return S; }
		m.world = Identity4();
		for (int i = 0; i < 3; ++i) m.world.d[3][i] += m.pos[i];     // Mesh.cpp:29 (the CPU path uses the Mesh's own matrix)
	}
	// camera: the call sequence of main.cpp:421-425 on the restated Camera (oracle_math.h; Camera.cpp)
	Camera cam(F.h.camPos, F.h.fovDeg);
	cam.SetAspectRatio(float(F.h.width), float(F.h.height));
	cam.SetNearFar(F.h.nearZ, F.h.farZ);
	if (F.h.pitchDeg != 0.f) cam.Pitch(F.h.pitchDeg);
	if (F.h.yawDeg != 0.f) cam.Yaw(F.h.yawDeg);
	F.proj = cam.proj;
	F.view = h_Inverse(cam.lookAt);                 // Camera.h:26
	F.viewProj = h_Mul(F.proj, F.view);             // ERenderer.cpp:54-56
	F.camPos = V3{ cam.lookAt.d[3][0], cam.lookAt.d[3][1], cam.lookAt.d[3][2] };
	F.camFwd = V3{ cam.lookAt.d[2][0], cam.lookAt.d[2][1], cam.lookAt.d[2][2] };
	F.back.resize((size_t)F.h.width * F.h.height);
	F.depth.resize((size_t)F.h.width * F.h.height);
	return S;
}
const char* ere_error(const EreScene* S) { return S->error.empty() ? nullptr : S->error.c_str(); }
void ere_free(EreScene* S) { delete S; }

// Renders one frame restricted to meshes [meshBegin, meshEnd) and triangles [triBegin, triEnd) of each;
// returns the seconds spent inside Render (steady_clock, scene load excluded).
double ere_render(EreScene* S, uint32_t meshBegin, uint32_t meshEnd, uint32_t triBegin, uint32_t triEnd)
{
	S->F.fragsShaded = 0; S->F.trisDrawn = 0;
	const auto t0 = std::chrono::steady_clock::now();
	RenderFrame(S->F, meshBegin, meshEnd, triBegin, triEnd);
	const auto t1 = std::chrono::steady_clock::now();
	return std::chrono::duration<double>(t1 - t0).count();
}
const uint32_t* ere_backbuffer(const EreScene* S) { return S->F.back.data(); }
const float* ere_depth(const EreScene* S) { return S->F.depth.data(); }
uint64_t ere_frags_shaded(const EreScene* S) { return S->F.fragsShaded; }
uint64_t ere_tris_drawn(const EreScene* S) { return S->F.trisDrawn; }

}
