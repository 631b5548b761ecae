// TEST INFRASTRUCTURE -- drop-in check: the reference's OWN host code drives the B200 renderer.
//
// oracle/build_ref_dropin.sh extracts `AddTexture` and `CreateScenes` from the reference's main.cpp (main.cpp:23-235,
// verbatim; the scene blocks the author left commented out - the two triangles, the uv quad, the TukTuk, the Vehicle -
// are switched on by removing their leading `//`, nothing else is edited) into a git-ignored scratch file and compiles
// it into this program against the product's C++ veneer (csrc/host/b2r_scene.hpp) instead of the reference's headers.
// The rest of this file is the call sequence of the reference's main() (main.cpp:391-470) without the SDL window.
// Output: oracle/_ref/dropin_main (a binary: it travels to the GPU box, the extracted source does not).
//
//   dropin_main <resources dir> list                        print every scene graph CreateScenes built (no GPU needed)
//   dropin_main <resources dir> render <graph> <w> <h> <camX> <camY> <camZ> <cull toggles> <sample toggles> <out.raw>
#include "b2r_scene.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <string>
#include <unistd.h>
#include <vector>

#include "create_scenes.inc"      // the reference's AddTexture + CreateScenes

int main(int argc, char** argv)
{
	if (argc < 3) { fprintf(stderr, "usage: %s <resources dir> list | render <graph> <w> <h> <cx> <cy> <cz> <cull> <sample> <out.raw>\n", argv[0]); return 2; }
	// CreateScenes opens "Resources/<file>" relative to the working directory (main.cpp:133,153,168)
	if (chdir(argv[1]) != 0) { perror("chdir"); return 2; }

	//Initialize framework (main.cpp:416-418)
	SceneManager sm{};
	CUDATextureManager tm{};
	CreateScenes(sm, tm);

	if (!strcmp(argv[2], "list"))
	{
		for (int g = 0; ; ++g)
		{
			const SceneGraph* pGraph = sm.GetSceneGraph();
			for (const Mesh* pMesh : pGraph->GetMeshes())
			{
				const int* t = pMesh->GetTextureIds();
				printf("graph %d mesh: %u vertices %u indices textures %d %d %d %d lights %zu\n", g, pMesh->GetNumVertices(), pMesh->GetNumIndices(),
					t[0], t[1], t[2], t[3], pGraph->GetLights().size());
			}
			sm.ChangeSceneGraph();
			if (sm.GetSceneGraph() == pGraph || g > 16) break;       // single graph, or safety stop
			if (g >= 4) break;                                       // the five graphs of the extracted source
		}
		return 0;
	}
	if (strcmp(argv[2], "render") != 0 || argc < 12) return 2;
	const int graph = atoi(argv[3]);
	for (int g = 0; g < graph; ++g) sm.ChangeSceneGraph();
	Resolution res{};
	res.Width = (unsigned)atoi(argv[4]); res.Height = (unsigned)atoi(argv[5]);
	for (int i = 0; i < atoi(argv[9]); ++i) sm.ToggleCullingMode();
	for (int i = 0; i < atoi(argv[10]); ++i) sm.ToggleSampleState();

	//Camera Setup (main.cpp:420-425)
	const FPoint3 camPos = { (float)atof(argv[6]), (float)atof(argv[7]), (float)atof(argv[8]) };
	const float fov = 45.f;
	Camera camera{ camPos, fov };
	camera.SetAspectRatio(float(res.Width), float(res.Height));

	std::vector<unsigned int> backBuffer((size_t)res.Width * res.Height);
	WindowHelper windowHelper{};                                    // main.cpp:429-434, no window
	windowHelper.pBackBufferPixels = backBuffer.data();
	windowHelper.Resolution = res;

	CUDARenderer* pCudaRenderer = new CUDARenderer{ windowHelper }; // main.cpp:440
	if (!pCudaRenderer->GetContext()) { fprintf(stderr, "dropin_main: %s\n", pCudaRenderer->GetLastError()); delete pCudaRenderer; return 3; }
	SceneGraph* pSceneGraph = sm.GetSceneGraph();                   // main.cpp:458-459
	pCudaRenderer->LoadScene(pSceneGraph, tm);
	pCudaRenderer->RenderAuto(sm, tm, &camera);                     // main.cpp:493
	if (pCudaRenderer->GetLastError()[0]) { fprintf(stderr, "dropin_main: %s\n", pCudaRenderer->GetLastError()); delete pCudaRenderer; return 3; }
	printf("triangles %u visible %u\n", pCudaRenderer->GetTotalNumTriangles(), pCudaRenderer->GetTotalNumVisibleTriangles());
	FILE* out = fopen(argv[11], "wb");
	if (!out) return 2;
	fwrite(backBuffer.data(), 4, backBuffer.size(), out);
	fclose(out);
	delete pCudaRenderer;
	return 0;
}
