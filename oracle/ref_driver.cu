// TEST INFRASTRUCTURE -- headless driver around the UNMODIFIED reference renderer.
//
// This file is ours (not reference code).  It replaces the reference's interactive
// main.cpp (CustomCUDARasterizer/main.cpp:391-536) with a file-in / file-out harness:
// it reads a scene dump (format: custom_gpgpu_rasterizer_b200/scenefile.py), builds the
// scene through the reference's own public classes (ObjParser/Mesh/SceneGraph/
// SceneManager/Camera/CUDATextureManager/CUDATexture), calls the reference's
// CUDARenderer and writes every parity surface to an output dump.
//
// Two modes:
//   render : SetupBins + LoadScene + RenderAuto, exactly the call sequence of
//            main.cpp:440-493.
//   stages : the same frame, but the per-mesh / per-batch loop of
//            CUDARenderer::Render (CUDARenderer.cu:237-330) is driven from here through
//            the renderer's own (private) launchers so the visible-triangle list and the
//            bin queues of every batch can be copied out between launches.  The kernels
//            and launchers executed are the reference's; only the loop is restated.
//   parse  : run the reference ObjParser on an .obj and dump its vertex/index buffers.
//
// Only runs on a GPU box (render/stages).  The product never links this.
#define private public
#define protected public
#include "PCH.h"
#include "CUDARenderer.h"
#include "ObjParser.h"
#include "DirectionalLight.h"
#undef private
#undef protected

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <string>
#include <vector>

// texture table consumed by the IMG_Load stub (oracle/sdl_stub.cpp)
struct StubTexture { unsigned w, h; const unsigned char* rgba; };
extern "C" void b2r_stub_register_texture(const char* name, unsigned w, unsigned h, const unsigned char* rgba);

namespace {

struct SceneHeader
{
	char magic[4];
	unsigned version;
	unsigned width, height;
	unsigned binsX, binsY, binW, binH, queueCap;
	unsigned cull, sample, vis;
	float camPos[3];
	float fovDeg, nearZ, farZ, pitchDeg, yawDeg;
	unsigned nTextures, nMeshes;
	unsigned flags;
};

template <typename T> bool ReadN(FILE* f, T* dst, size_t n) { return fread(dst, sizeof(T), n, f) == n; }
template <typename T> void WriteN(FILE* f, const T* src, size_t n) { fwrite(src, sizeof(T), n, f); }

int Fail(const char* msg) { fprintf(stderr, "ref_render: %s\n", msg); return 2; }

int Parse(const char* objPath, int invertY, const char* outPath, bool useVectorOverload)
{
	ObjParser parser{};
	if (!parser.OpenFile(objPath))
		return Fail("cannot open obj");
	parser.SetInvertYAxis(invertY != 0);
	IVertex* pV{}; unsigned nV{}; unsigned* pI{}; unsigned nI{}; short vertexType{};
	if (useVectorOverload)
	{
		// ObjParser.cpp:100-108 / 381-548: the std::vector overload (the raw-pointer one sizes its buffer by the
		// face count and overruns it for meshes with more unique vertices than faces, e.g. vehicle.obj)
		std::vector<IVertex> v; std::vector<unsigned int> idx;
		parser.ReadFromObjFile(v, idx, vertexType);
		nV = (unsigned)v.size(); nI = (unsigned)idx.size();
		pV = new IVertex[nV]; pI = new unsigned[nI];
		memcpy((void*)pV, v.data(), sizeof(IVertex) * nV); memcpy(pI, idx.data(), 4 * (size_t)nI);
	}
	else
		parser.ReadFromObjFile(pV, nV, pI, nI, vertexType);
	FILE* f = fopen(outPath, "wb");
	if (!f) return Fail("cannot open output");
	const unsigned hdr[4] = { 0x50523242u /*B2RP*/, nV, nI, (unsigned)(unsigned short)vertexType };
	WriteN(f, hdr, 4);
	WriteN(f, reinterpret_cast<const unsigned char*>(pV), sizeof(IVertex) * nV);
	WriteN(f, pI, nI);
	fclose(f);
	delete[] pV; delete[] pI;
	return 0;
}

}

int main(int argc, char** argv)
{
	if (argc >= 5 && (std::string(argv[1]) == "parse" || std::string(argv[1]) == "parsev"))
		return Parse(argv[2], atoi(argv[3]), argv[4], std::string(argv[1]) == "parsev");
	if (argc < 4)
		return Fail("usage: ref_render render|stages scene.b2rs out.b2ro | time scene.b2rs frames | parse file.obj invertY out.bin");
	const bool stages = std::string(argv[1]) == "stages";
	const bool timing = std::string(argv[1]) == "time";

	FILE* fin = fopen(argv[2], "rb");
	if (!fin) return Fail("cannot open scene");
	SceneHeader h{};
	if (!ReadN(fin, &h, 1) || memcmp(h.magic, "B2RS", 4) != 0 || h.version != 1)
		return Fail("bad scene header");

	CheckErrorCuda(SetDeviceCuda(0));

	SceneManager sm{};
	CUDATextureManager tm{};

	// textures: registered with the IMG_Load stub under "mem:<i>", then created through
	// the reference's own CUDATexture::Create (CUDATexture.cpp:24-72)
	std::vector<std::vector<unsigned char>> texData(h.nTextures);
	for (unsigned i = 0; i < h.nTextures; ++i)
	{
		unsigned wh[2];
		if (!ReadN(fin, wh, 2)) return Fail("bad texture header");
		texData[i].resize(size_t(wh[0]) * wh[1] * 4);
		if (!ReadN(fin, texData[i].data(), texData[i].size())) return Fail("bad texture data");
		const std::string name = "mem:" + std::to_string(i);
		b2r_stub_register_texture(name.c_str(), wh[0], wh[1], texData[i].data());
		CUDATexture* pTex = new CUDATexture{ name };
		if (!pTex->IsAllocated()) return Fail("texture create failed");
		tm.AddCUDATexture(pTex);
	}

	SceneGraph* pSceneGraph = new SceneGraph{};
	for (unsigned m = 0; m < h.nMeshes; ++m)
	{
		unsigned mh[3]; float pos[3]; int tex[4];
		if (!ReadN(fin, mh, 3) || !ReadN(fin, pos, 3) || !ReadN(fin, tex, 4)) return Fail("bad mesh header");
		IVertex* pV = new IVertex[mh[0]];
		unsigned* pI = new unsigned[mh[1]];
		if (!ReadN(fin, reinterpret_cast<unsigned char*>(pV), sizeof(IVertex) * mh[0]) || !ReadN(fin, pI, mh[1]))
			return Fail("bad mesh data");
		Mesh* pMesh = new Mesh{ pV, mh[0], (short)sizeof(IVertex), (short)0, pI, mh[1],
			(PrimitiveTopology)mh[2], FPoint3{ pos[0], pos[1], pos[2] } };
		pMesh->SetTextureIds(tex[0], tex[1], tex[2], tex[3]);
		pSceneGraph->AddMesh(pMesh);
	}
	fclose(fin);
	pSceneGraph->AddLight(new DirectionalLight{ RGBColor{1.f, 1.f, 1.f}, 2.f, FVector3{ 0.577f, -0.577f, -0.577f } });
	sm.AddSceneGraph(pSceneGraph);

	// the only public way to set the three switches is the key-toggle methods
	for (unsigned i = 0; i < h.cull; ++i) sm.ToggleCullingMode();
	for (unsigned i = 0; i < h.sample; ++i) sm.ToggleSampleState();
	for (unsigned i = 0; i < h.vis; ++i) sm.ToggleVisualisation();

	Camera camera{ FPoint3{ h.camPos[0], h.camPos[1], h.camPos[2] }, h.fovDeg };
	camera.SetAspectRatio(float(h.width), float(h.height));
	camera.SetNearFarValues(h.nearZ, h.farZ);
	if (h.pitchDeg != 0.f) camera.Pitch(h.pitchDeg);
	if (h.yawDeg != 0.f) camera.Yaw(h.yawDeg);

	std::vector<unsigned> backBuffer(size_t(h.width) * h.height, 0u);
	WindowHelper windowHelper{};
	windowHelper.pBackBufferPixels = backBuffer.data();
	windowHelper.Resolution.Width = h.width;
	windowHelper.Resolution.Height = h.height;

	CUDARenderer* pRenderer = new CUDARenderer{ windowHelper };
	pRenderer->SetupBins(IPoint2{ (int)h.binsX, (int)h.binsY }, IPoint2{ (int)h.binW, (int)h.binH }, h.queueCap);
	// the reference never initialises the queue size/mutex words (CUDAAtomicQueue.cuh:150-158)
	pRenderer->m_BinQueues.ResetQueueSizes();
	cudaMemset(pRenderer->m_BinQueues.GetDev_QueueSizesMutexBuffer(), 0, size_t(h.binsX) * h.binsY * 4);
	pRenderer->LoadScene(pSceneGraph, tm);

	const size_t nPix = size_t(h.width) * h.height;
	std::vector<unsigned> visPerMesh;
	std::vector<std::vector<unsigned>> visTriples(h.nMeshes);   // idx0,idx1,idx2 per visible triangle
	std::vector<std::vector<unsigned>> binPairs(h.nMeshes);     // bin,idx0,idx1,idx2 per queue entry

	if (timing)
	{
		// wall clock of the reference's own frame call (RenderAuto = clear + Render + Present, CUDARenderer.cu:357-371),
		// one warm-up frame, then argv[3] frames; the reference synchronises the device inside every frame
		const int frames = atoi(argv[3]) > 0 ? atoi(argv[3]) : 1;
		pRenderer->RenderAuto(sm, tm, &camera);
		cudaDeviceSynchronize();
		const auto t0 = std::chrono::steady_clock::now();
		for (int f = 0; f < frames; ++f) pRenderer->RenderAuto(sm, tm, &camera);
		const cudaError_t terr = cudaDeviceSynchronize();
		const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		if (terr != cudaSuccess)
		{
			fprintf(stderr, "ref_render: CUDA error while timing: %s\n", cudaGetErrorString(terr));
			return 3;
		}
		printf("{\"frames\": %d, \"ms_per_frame\": %.6f, \"triangles\": %u, \"visible_triangles\": %u}\n", frames, sec * 1e3 / frames,
			pRenderer->GetTotalNumTriangles(), pRenderer->GetTotalNumVisibleTriangles());
		return 0;
	}
	if (!stages)
	{
		pRenderer->RenderAuto(sm, tm, &camera);
	}
	else
	{
		// restated loop of CUDARenderer::Render (CUDARenderer.cu:207-354) + RenderAuto (:357-371)
		pRenderer->EnterValidRenderingState();
		const VisualisationState visualisationState = sm.GetVisualisationState();
		const SampleState sampleState = sm.GetSampleState();
		const CullingMode cm = sm.GetCullingMode();
		const FPoint3& camPos = camera.GetPosition();
		const FVector3& camFwd = camera.GetForward();
		const FMatrix4 viewMatrix = camera.GetViewMatrix();
		const FMatrix4& projectionMatrix = camera.GetProjectionMatrix();
		const FMatrix4 viewProjectionMatrix = projectionMatrix * viewMatrix;
		pRenderer->UpdateCameraDataAsync(camPos, camFwd);
		pRenderer->m_TotalVisibleNumTriangles = 0;
		const unsigned nBins = h.binsX * h.binsY;
		std::vector<unsigned> qSizes(nBins), qData(size_t(nBins) * h.queueCap);
		unsigned meshNo = 0;
		for (auto sortedCudaMesh : pRenderer->m_CUDAMeshes)
		{
			const CUDAMesh* pCudaMesh = sortedCudaMesh.pCUDAMesh;
			const FMatrix4& worldMat = pCudaMesh->GetWorldConst();
			const FMatrix4 worldViewProjectionMatrix = viewProjectionMatrix * worldMat;
			const FMatrix3& rotationMatrix = pCudaMesh->GetRotationMatrix();
			pRenderer->UpdateWorldMatrixDataAsync(worldMat, worldViewProjectionMatrix, rotationMatrix);
			cudaDeviceSynchronize();
			pRenderer->VertexShader(pCudaMesh);
			cudaDeviceSynchronize();
			cudaMemset(pRenderer->m_Dev_NumVisibleTriangles, 0, sizeof(unsigned));
			pRenderer->TriangleAssembler(pCudaMesh, camFwd, cm);
			cudaDeviceSynchronize();
			unsigned numVisible = 0;
			cudaMemcpy(&numVisible, pRenderer->m_Dev_NumVisibleTriangles, 4, cudaMemcpyDeviceToHost);
			pRenderer->m_TotalVisibleNumTriangles += numVisible;
			visPerMesh.push_back(numVisible);
			std::vector<unsigned>& tri = visTriples[meshNo];
			tri.resize(size_t(numVisible) * 3);
			if (numVisible)
				cudaMemcpy(tri.data(), pCudaMesh->GetDevTriangleBuffer(), tri.size() * 4, cudaMemcpyDeviceToHost);
			const unsigned queueMaxSize = pRenderer->m_BinQueues.GetQueueMaxSize();
			const unsigned numLoops = (numVisible + (queueMaxSize - 1)) / queueMaxSize;
			for (unsigned i = 0; i < numLoops; ++i)
			{
				pRenderer->TriangleBinner(pCudaMesh, numVisible, i * queueMaxSize);
				cudaDeviceSynchronize();
				cudaMemcpy(qSizes.data(), pRenderer->m_BinQueues.GetDev_QueueSizes(), nBins * 4, cudaMemcpyDeviceToHost);
				cudaMemcpy(qData.data(), pRenderer->m_BinQueues.GetDev_Queues(), qData.size() * 4, cudaMemcpyDeviceToHost);
				for (unsigned b = 0; b < nBins; ++b)
					for (unsigned k = 0; k < qSizes[b]; ++k)
					{
						const unsigned slot = qData[size_t(b) * h.queueCap + k];
						binPairs[meshNo].push_back(b);
						binPairs[meshNo].push_back(tri[size_t(slot) * 3 + 0]);
						binPairs[meshNo].push_back(tri[size_t(slot) * 3 + 1]);
						binPairs[meshNo].push_back(tri[size_t(slot) * 3 + 2]);
					}
				pRenderer->Rasterizer(pCudaMesh, camFwd, cm);
				cudaDeviceSynchronize();
				pRenderer->m_BinQueues.ResetQueueSizes();
			}
			++meshNo;
		}
		pRenderer->PixelShader(sampleState, visualisationState);
		cudaDeviceSynchronize();
		pRenderer->Present();
	}
	cudaError_t err = cudaDeviceSynchronize();
	if (err != cudaSuccess)
	{
		fprintf(stderr, "ref_render: CUDA error after frame: %s\n", cudaGetErrorString(err));
		return 3;
	}

	// parity surfaces
	std::vector<int> depthInt(nPix);
	cudaMemcpy(depthInt.data(), pRenderer->m_CUDAWindowHelper.GetDev_DepthBuffer(), nPix * 4, cudaMemcpyDeviceToHost);
	std::vector<PixelShade> frags(nPix);
	cudaMemcpy(frags.data(), pRenderer->m_CUDAWindowHelper.GetDev_FragmentBuffer(), nPix * sizeof(PixelShade), cudaMemcpyDeviceToHost);
	std::vector<float> zInterp(nPix), fragAttr(nPix * 13);
	for (size_t p = 0; p < nPix; ++p)
	{
		zInterp[p] = frags[p].zInterpolated;
		float* a = &fragAttr[p * 13];
		a[0] = frags[p].wInterpolated;
		a[1] = frags[p].uv.x; a[2] = frags[p].uv.y;
		a[3] = frags[p].n.x; a[4] = frags[p].n.y; a[5] = frags[p].n.z;
		a[6] = frags[p].tan.x; a[7] = frags[p].tan.y; a[8] = frags[p].tan.z;
		a[9] = frags[p].vd.x; a[10] = frags[p].vd.y; a[11] = frags[p].vd.z;
		memcpy(&a[12], &frags[p].colour32, 4);
	}

	FILE* fout = fopen(argv[3], "wb");
	if (!fout) return Fail("cannot open output");
	const FMatrix4 view = camera.GetViewMatrix();
	const FMatrix4 proj = camera.GetProjectionMatrix();
	const FMatrix4 viewProj = proj * view;
	const unsigned oh[8] = { 0x4F523242u /*B2RO*/, 1u, h.width, h.height,
		pRenderer->GetTotalNumTriangles(), pRenderer->GetTotalNumVisibleTriangles(), h.nMeshes, stages ? 1u : 0u };
	WriteN(fout, oh, 8);
	WriteN(fout, &viewProj.data[0][0], 16);
	WriteN(fout, &view.data[0][0], 16);
	WriteN(fout, &proj.data[0][0], 16);
	WriteN(fout, camera.GetPosition().data, 3);
	WriteN(fout, camera.GetForward().data, 3);
	WriteN(fout, backBuffer.data(), nPix);
	WriteN(fout, depthInt.data(), nPix);
	WriteN(fout, zInterp.data(), nPix);
	WriteN(fout, fragAttr.data(), nPix * 13);
	if (stages)
	{
		for (unsigned m = 0; m < h.nMeshes; ++m)
		{
			const unsigned cnt[2] = { visPerMesh[m], (unsigned)(binPairs[m].size() / 4) };
			WriteN(fout, cnt, 2);
			WriteN(fout, visTriples[m].data(), visTriples[m].size());
			WriteN(fout, binPairs[m].data(), binPairs[m].size());
			// OVertex buffer of the mesh (18 floats per vertex)
			const CUDAMesh* pCudaMesh = pRenderer->m_CUDAMeshes[m].pCUDAMesh;
			const unsigned nV = pCudaMesh->GetNumVertices();
			std::vector<OVertex> ov(nV);
			cudaMemcpy((void*)ov.data(), pCudaMesh->GetDevOVertexBuffer(), sizeof(OVertex) * nV, cudaMemcpyDeviceToHost);
			WriteN(fout, &nV, 1);
			WriteN(fout, reinterpret_cast<const unsigned char*>(ov.data()), sizeof(OVertex) * nV);
		}
	}
	fclose(fout);
	delete pRenderer;
	return 0;
}
