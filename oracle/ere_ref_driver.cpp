// TEST / MEASUREMENT INFRASTRUCTURE (oracle) -- headless driver of the reference's OWN CPU renderer.
//
// oracle/build_ref_erenderer.sh compiles the reference's Elite::Renderer (CustomCUDARasterizer/ERenderer.cpp, unmodified)
// together with its Camera / Light / DirectionalLight / SceneGraph translation units and its math, texture-sampler and
// vertex headers, against OUR reconstruction of the two interfaces that file predates (oracle/ere_shim/: legacy_mesh_interface.h,
// legacy_scene_manager_interface.h, erenderer_prelude.h - installed as Mesh.h / SceneManager.h / PCH.h of the build) and OUR SDL stand-ins (oracle/sdl_stub.cpp, sdl_stub_ere.cpp), into oracle/_ref/libere_ref.so.
// This file is the glue: it exports the same C interface as the restatement (oracle/erenderer_oracle.cpp: ere_load /
// ere_render / ere_backbuffer / ere_depth / ere_free over a scene dump of custom_gpgpu_rasterizer_b200/scenefile.py),
// so tests/ can compare the two frame by frame and bench.py can time the real thing (cpu_baseline.kind "reference").
// Scene construction follows the reference's main.cpp:136-178 / 421-425 (Camera, SetAspectRatio, one DirectionalLight).
// Only tests/, tools/ and bench.py's CPU-baseline legs load this library; the product never does.
#include "PCH.h"
#include <SDL.h>

#define private public          // the driver reads Elite::Renderer's back buffer and depth buffer (test infrastructure)
#include "ERenderer.h"
#undef private
#include "Camera.h"
#include "DirectionalLight.h"
#include "Mesh.h"
#include "SceneGraph.h"
#include "SceneManager.h"
#include "Texture.h"

#include <chrono>
#include <string>

extern "C" void b2r_stub_register_texture(const char* name, unsigned w, unsigned h, const unsigned char* rgba);
extern "C" void b2r_stub_set_window_size(int w, int h);

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
}

struct EreScene
{
	SceneHeader h{};
	std::vector<std::vector<uint32_t>> pixels;
	std::vector<Texture*> textures;
	SceneManager sm;
	Camera* pCamera = nullptr;
	Elite::Renderer* pRenderer = nullptr;
	std::string error;
};

extern "C" {

EreScene* ere_load(const uint8_t* blob, uint64_t size)
{
	EreScene* S = new EreScene;
	size_t off = 0;
	auto take = [&](void* dst, size_t n) { if (off + n > size) return false; memcpy(dst, blob + off, n); off += n; return true; };
	SceneHeader& h = S->h;
	if (!take(&h, sizeof(SceneHeader)) || memcmp(h.magic, "B2RS", 4) != 0) { S->error = "bad header"; return S; }
	static int serial = 0;
	S->pixels.resize(h.nTextures);
	for (uint32_t i = 0; i < h.nTextures; ++i)
	{
		uint32_t wh[2];
		if (!take(wh, 8)) { S->error = "bad texture"; return S; }
		S->pixels[i].resize((size_t)wh[0] * wh[1]);
		if (!take(S->pixels[i].data(), S->pixels[i].size() * 4)) { S->error = "bad texture"; return S; }
		const std::string name = "mem:ere" + std::to_string(serial++);
		b2r_stub_register_texture(name.c_str(), wh[0], wh[1], reinterpret_cast<const unsigned char*>(S->pixels[i].data()));
		S->textures.push_back(new Texture{ name.c_str() });      // Texture.cpp:8-13 -> IMG_Load (stub: copies the RGBA8 image)
	}
	SceneGraph* pSceneGraph = new SceneGraph{};
	for (uint32_t m = 0; m < h.nMeshes; ++m)
	{
		uint32_t mh[3]; float pos[3]; int tex[4];
		if (!take(mh, 12) || !take(pos, 12) || !take(tex, 16)) { S->error = "bad mesh"; delete pSceneGraph; return S; }
		std::vector<IVertex> v(mh[0]); std::vector<unsigned int> idx(mh[1]);
		if (!take((void*)v.data(), sizeof(IVertex) * mh[0]) || !take(idx.data(), 4 * (size_t)mh[1])) { S->error = "bad mesh"; delete pSceneGraph; return S; }
		auto T = [&](int id) -> Texture* { return (id >= 0 && (size_t)id < S->textures.size()) ? S->textures[id] : nullptr; };
		Textures t; t.pDiff = T(tex[0]); t.pNorm = T(tex[1]); t.pSpec = T(tex[2]); t.pGloss = T(tex[3]);
		pSceneGraph->AddMesh(new Mesh{ std::move(v), std::move(idx), (PrimitiveTopology)mh[2], FPoint3{ pos[0], pos[1], pos[2] }, t });
	}
	pSceneGraph->AddLight(new DirectionalLight{ RGBColor{ 1.f, 1.f, 1.f }, 2.f, FVector3{ 0.577f, -0.577f, -0.577f } });   // main.cpp:178
	S->sm.AddSceneGraph(pSceneGraph);
	S->sm.SetCullingMode((CullingMode)h.cull);
	S->sm.SetSampleState((SampleState)h.sample);
	S->sm.SetVisualisationState((VisualisationState)h.vis);

	S->pCamera = new Camera{ FPoint3{ h.camPos[0], h.camPos[1], h.camPos[2] }, h.fovDeg };     // main.cpp:421-425
	S->pCamera->SetAspectRatio(float(h.width), float(h.height));
	S->pCamera->SetNearFarValues(h.nearZ, h.farZ);
	if (h.pitchDeg != 0.f) S->pCamera->Pitch(h.pitchDeg);
	if (h.yawDeg != 0.f) S->pCamera->Yaw(h.yawDeg);

	b2r_stub_set_window_size((int)h.width, (int)h.height);
	S->pRenderer = new Elite::Renderer{ nullptr };                // ERenderer.cpp:22-34: sizes and surfaces come from the SDL stand-ins
	S->pRenderer->SetCamera(S->pCamera);
	return S;
}

const char* ere_error(const EreScene* S) { return S->error.empty() ? nullptr : S->error.c_str(); }

void ere_free(EreScene* S)
{
	if (!S) return;
	delete S->pRenderer;
	delete S->pCamera;
	delete S->sm.GetSceneGraph();          // deletes its meshes and lights (SceneGraph.cpp:6-18)
	for (Texture* t : S->textures) delete t;
	delete S;
}

// One frame through Elite::Renderer::Render (ERenderer.cpp:42-164); returns the seconds spent inside it.  The sub-range
// arguments of the restatement's ere_render are not supported by the real renderer: anything but "everything" is -1.
double ere_render(EreScene* S, uint32_t meshBegin, uint32_t meshEnd, uint32_t triBegin, uint32_t triEnd)
{
	if (!S->pRenderer || !S->error.empty()) return -1.0;
	if (meshBegin != 0 || triBegin != 0 || meshEnd < S->h.nMeshes || triEnd != 0xFFFFFFFFu) return -1.0;
	const auto t0 = std::chrono::steady_clock::now();
	S->pRenderer->Render(S->sm);
	const auto t1 = std::chrono::steady_clock::now();
	return std::chrono::duration<double>(t1 - t0).count();
}

const uint32_t* ere_backbuffer(const EreScene* S) { return S->pRenderer ? S->pRenderer->m_pBackBufferPixels : nullptr; }
const float* ere_depth(const EreScene* S) { return S->pRenderer ? S->pRenderer->m_pDepthBuffer : nullptr; }
uint64_t ere_frags_shaded(const EreScene*) { return 0; }      // the reference keeps no counters
uint64_t ere_tris_drawn(const EreScene*) { return 0; }
int ere_is_reference(void) { return 1; }

// The reference's Camera (Camera.h / Camera.cpp, compiled as they lie) through the call sequence of main.cpp:421-425 and
// SURVEY's config 5 (Pitch, then Yaw): the matrices and vectors the renderers take from it.  Lets tests/ pin the product's
// host-side Camera (csrc/host/b2r_scene.cpp, b2r_host_camera) and the oracles' restatement against the real class on CPU.
void ere_reference_camera(const float pos[3], float fovDeg, float width, float height, float nearZ, float farZ, float pitchDeg, float yawDeg,
	float* viewProj16, float* view16, float* proj16, float* camPos3, float* camFwd3)
{
	Camera cam{ FPoint3{ pos[0], pos[1], pos[2] }, fovDeg };
	cam.SetAspectRatio(width, height);
	cam.SetNearFarValues(nearZ, farZ);
	if (pitchDeg != 0.f) cam.Pitch(pitchDeg);
	if (yawDeg != 0.f) cam.Yaw(yawDeg);
	const FMatrix4 view = cam.GetViewMatrix();
	const FMatrix4 proj = cam.GetProjectionMatrix();
	const FMatrix4 vp = proj * view;                   // CUDARenderer.cu:219, ERenderer.cpp:54-56
	memcpy(viewProj16, vp.data, 64); memcpy(view16, view.data, 64); memcpy(proj16, proj.data, 64);
	camPos3[0] = cam.GetPosition().x; camPos3[1] = cam.GetPosition().y; camPos3[2] = cam.GetPosition().z;
	camFwd3[0] = cam.GetForward().x; camFwd3[1] = cam.GetForward().y; camFwd3[2] = cam.GetForward().z;
}

}
