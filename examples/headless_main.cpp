// Headless replacement of the reference's main.cpp (CustomCUDARasterizer/main.cpp:391-536) written against
// the C++ veneer: the same classes and the same call sequence (scene creation, SetupBins, LoadScene,
// RenderAuto), no SDL window; the back buffer is dumped to a file.
//
//   headless_main scene.b2rs out.raw      render a scene dump (format: custom_gpgpu_rasterizer_b200/scenefile.py)
//   headless_main mesh.obj out.raw [w h]  load an OBJ like main.cpp:148-160 and render it at 640x480
//
// build: g++ -std=c++17 -O2 -ffp-contract=off examples/headless_main.cpp -Icustom_gpgpu_rasterizer_b200/csrc/host
//            -Lcustom_gpgpu_rasterizer_b200/csrc -lb2r -Wl,-rpath,$PWD/custom_gpgpu_rasterizer_b200/csrc -o headless_main
#include "b2r_scene.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {
struct SceneHeader
{
	char magic[4]; unsigned version, width, height, binsX, binsY, binW, binH, queueCap, cull, sample, vis;
	float camPos[3], fovDeg, nearZ, farZ, pitchDeg, yawDeg; unsigned nTextures, nMeshes, flags;
};
}

int main(int argc, char** argv)
{
	if (argc < 3) { fprintf(stderr, "usage: %s scene.b2rs|mesh.obj out.raw [w h]\n", argv[0]); return 2; }
	const std::string in = argv[1];
	SceneManager sm{};
	CUDATextureManager tm{};
	SceneGraph* pSceneGraph = new SceneGraph{};
	SceneHeader h{};
	if (in.size() > 4 && in.substr(in.size() - 4) == ".obj")
	{
		// main.cpp:148-160 (bunny scene) with the camera of main.cpp:421-425
		h.width = argc > 4 ? atoi(argv[3]) : 640; h.height = argc > 4 ? atoi(argv[4]) : 480;
		h.binW = h.binH = 16; h.binsX = (h.width + 15) / 16; h.binsY = (h.height + 15) / 16; h.queueCap = 256;
		h.camPos[0] = 0.f; h.camPos[1] = 1.f; h.camPos[2] = 5.f; h.fovDeg = 45.f; h.nearZ = 0.1f; h.farZ = 100.f;
		ObjParser parser{};
		if (!parser.OpenFile(in)) { fprintf(stderr, "cannot open %s\n", in.c_str()); return 2; }
		IVertex* pVertexBuffer{}; unsigned int numVertices{}; unsigned int* pIndexBuffer{}; unsigned int numIndices{}; short vertexType{};
		parser.ReadFromObjFile(pVertexBuffer, numVertices, pIndexBuffer, numIndices, vertexType);
		pSceneGraph->AddMesh(new Mesh{ pVertexBuffer, numVertices, (short)sizeof(IVertex), vertexType, pIndexBuffer, numIndices, PrimitiveTopology::TriangleList });
	}
	else
	{
		FILE* f = fopen(in.c_str(), "rb");
		if (!f || fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "B2RS", 4) != 0) { fprintf(stderr, "bad scene file\n"); return 2; }
		for (unsigned i = 0; i < h.nTextures; ++i)
		{
			unsigned wh[2]; if (fread(wh, 4, 2, f) != 2) return 2;
			std::vector<unsigned int> px((size_t)wh[0] * wh[1]);
			if (fread(px.data(), 4, px.size(), f) != px.size()) return 2;
			CUDATexture* pTex = new CUDATexture{};
			pTex->Create(px.data(), wh[0], wh[1]);
			tm.AddCUDATexture(pTex);
		}
		for (unsigned m = 0; m < h.nMeshes; ++m)
		{
			unsigned mh[3]; float pos[3]; int tex[4];
			if (fread(mh, 4, 3, f) != 3 || fread(pos, 4, 3, f) != 3 || fread(tex, 4, 4, f) != 4) return 2;
			IVertex* pV = new IVertex[mh[0]]; unsigned int* pI = new unsigned int[mh[1]];
			if (fread((void*)pV, sizeof(IVertex), mh[0], f) != mh[0] || fread(pI, 4, mh[1], f) != mh[1]) return 2;
			Mesh* pMesh = new Mesh{ pV, mh[0], (short)sizeof(IVertex), (short)0, pI, mh[1], (PrimitiveTopology)mh[2], FPoint3{ pos[0], pos[1], pos[2] } };
			pMesh->SetTextureIds(tex[0], tex[1], tex[2], tex[3]);
			pSceneGraph->AddMesh(pMesh);
		}
		fclose(f);
	}
	pSceneGraph->AddLight(new DirectionalLight{ RGBColor{ 1.f, 1.f, 1.f }, 2.f, FVector3{ 0.577f, -0.577f, -0.577f } });
	sm.AddSceneGraph(pSceneGraph);
	for (unsigned i = 0; i < h.cull; ++i) sm.ToggleCullingMode();
	for (unsigned i = 0; i < h.sample; ++i) sm.ToggleSampleState();
	for (unsigned i = 0; i < h.vis; ++i) sm.ToggleVisualisation();

	Camera camera{ FPoint3{ h.camPos[0], h.camPos[1], h.camPos[2] }, h.fovDeg };
	camera.SetAspectRatio(float(h.width), float(h.height));
	camera.SetNearFarValues(h.nearZ, h.farZ);
	if (h.pitchDeg != 0.f) camera.Pitch(h.pitchDeg);
	if (h.yawDeg != 0.f) camera.Yaw(h.yawDeg);

	std::vector<unsigned int> backBuffer((size_t)h.width * h.height);
	WindowHelper windowHelper{};
	windowHelper.pBackBufferPixels = backBuffer.data();
	windowHelper.Resolution.Width = h.width; windowHelper.Resolution.Height = h.height;

	CUDARenderer* pCudaRenderer = new CUDARenderer{ windowHelper };
	// the veneer keeps the reference's convention (errors are printed, never thrown); a headless tool must not write a
	// frame nobody rendered: no CUDA device (there is no CPU fallback) or a failed call ends the program
	if (!pCudaRenderer->GetContext()) { fprintf(stderr, "headless_main: %s\n", pCudaRenderer->GetLastError()); delete pCudaRenderer; return 3; }
	pCudaRenderer->SetupBins(IPoint2{ (int)h.binsX, (int)h.binsY }, IPoint2{ (int)h.binW, (int)h.binH }, h.queueCap);
	pCudaRenderer->LoadScene(sm.GetSceneGraph(), tm);
	pCudaRenderer->RenderAuto(sm, tm, &camera);
	if (pCudaRenderer->GetLastError()[0]) { fprintf(stderr, "headless_main: %s\n", pCudaRenderer->GetLastError()); delete pCudaRenderer; return 3; }
	printf("triangles %u visible %u\n", pCudaRenderer->GetTotalNumTriangles(), pCudaRenderer->GetTotalNumVisibleTriangles());

	FILE* out = fopen(argv[2], "wb");
	if (!out) return 2;
	fwrite(backBuffer.data(), 4, backBuffer.size(), out);
	fclose(out);
	// optional: --move dx dy dz  ->  the first mesh is moved the way an Update would move it (its world matrix), the scene
	// manager is updated (SceneGraph::Update with elapsedSec 0, as in the reference's main loop) and the next frame goes to
	// <out>.moved: the renderer picks the new matrix up by itself
	for (int a = 3; a + 3 < argc; ++a)
		if (std::string(argv[a]) == "--move" && !sm.GetSceneGraph()->GetMeshes().empty())
		{
			FMatrix4& world = sm.GetSceneGraph()->GetMeshes()[0]->GetWorld();
			for (int k = 0; k < 3; ++k) world.data[3][k] += (float)atof(argv[a + 1 + k]);
			sm.Update(0.f);
			pCudaRenderer->RenderAuto(sm, tm, &camera);
			printf("moved: triangles %u visible %u\n", pCudaRenderer->GetTotalNumTriangles(), pCudaRenderer->GetTotalNumVisibleTriangles());
			FILE* o2 = fopen((std::string(argv[2]) + ".moved").c_str(), "wb");
			if (!o2) return 2;
			fwrite(backBuffer.data(), 4, backBuffer.size(), o2);
			fclose(o2);
		}
	// optional: --texture file.png global stretched sample  ->  CUDARenderer::DrawTexture / DrawTextureGlobal into <out>.tex
	for (int a = 3; a + 4 < argc; ++a)
		if (std::string(argv[a]) == "--texture")
		{
			if (atoi(argv[a + 2])) pCudaRenderer->DrawTextureGlobal(argv[a + 1], atoi(argv[a + 3]) != 0, (SampleState)atoi(argv[a + 4]));
			else pCudaRenderer->DrawTexture(argv[a + 1]);
			FILE* o3 = fopen((std::string(argv[2]) + ".tex").c_str(), "wb");
			if (!o3) return 2;
			fwrite(backBuffer.data(), 4, backBuffer.size(), o3);
			fclose(o3);
		}
	delete pCudaRenderer;
	return 0;
}
