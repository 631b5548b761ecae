// C++ veneer: the reference's scene-model / renderer class surface, re-implemented on
// top of the C ABI in include/b2r.h (libb2r.so).  A program written against the reference's
// headers (Camera.h, Mesh.h, SceneGraph.h, SceneManager.h, ObjParser.h, CUDATexture.h,
// CUDATextureManager.h, CUDARenderer.h, the render-state enums) compiles against this
// header with the same class names, signatures, ownership and error conventions
// (SURVEY.md section 8b).  The SDL window is optional: WindowHelper may carry null SDL pointers
// and a plain host pixel buffer, and Present() is then only the device->host copy.
#pragma once
#include "b2r_math.hpp"

#include <fstream>
#include <string>
#include <vector>

struct b2r_context;
struct SDL_Window;
struct SDL_Surface;

// ---- render-state enums (CullingMode.h, SampleState.h, VisualisationState.h, PrimitiveTopology.h, VertexType.h)
enum class CullingMode { BackFace, FrontFace, NoCulling };
enum class SampleState { Point, Linear };
enum class VisualisationState { PBR, Depth, Normal };
enum class PrimitiveTopology : unsigned char { TriangleList, TriangleStrip };
enum class VertexType : short { Pos = 0, Uv = 1, Col = 2, Norm = 4, Tan = 8, Vd = 16 };

// ---- vertex layouts (Vertex.h:113-145, 280-327)
struct IVertex
{
	FPoint3 p{ 0.f, 0.f, 0.f };
	FVector2 uv;
	FVector3 n;
	FVector3 tan;
	RGBColor c;
	IVertex() = default;
	IVertex(const FPoint3& position, const RGBColor& colour = RGBColor{ 1.f, 1.f, 1.f }) : p(position), c(colour) {}
	IVertex(const FPoint3& position, const FVector2& uvcoordinates, const FVector3& normal = FVector3{ 1.f, 1.f, 1.f },
		const RGBColor& colour = RGBColor{ 1.f, 1.f, 1.f }) : p(position), uv(uvcoordinates), n(normal), c(colour) {}
};
static_assert(sizeof(IVertex) == 56, "IVertex is the 56-byte wire format of the vertex stream");

struct OVertex
{
	FPoint4 p;
	FVector2 uv;
	FVector3 n;
	FVector3 tan;
	FVector3 vd;
	RGBColor c;
};
static_assert(sizeof(OVertex) == 72, "OVertex is the 72-byte transformed-vertex format");

// ---- Resolution.h / WindowHelper.h
struct Resolution
{
	struct UInt2 { unsigned int w, h; };
	enum ResolutionStandard { VGA = 0, SD = 1, HD = 2, FHD = 3, QHD = 4, UHD = 5 };
	unsigned int Width = 0;
	unsigned int Height = 0;
	UInt2 AspectRatio{ 0, 0 };
	ResolutionStandard Standard = SD;
	static Resolution GetResolution(ResolutionStandard standard);
};

struct WindowHelper
{
	SDL_Window* pWindow = nullptr;
	SDL_Surface* pFrontBuffer = nullptr;
	SDL_Surface* pBackBuffer = nullptr;
	unsigned int* pBackBufferPixels = nullptr;
	::Resolution Resolution;
};

// ---- Camera.h
class Camera
{
public:
	explicit Camera(const FPoint3& position, float fov = 45.f);
	virtual ~Camera() = default;
	Camera(const Camera&) = delete;
	Camera& operator=(const Camera&) = delete;

	void SetPosition(const FPoint3& pos) { m_Position = pos; }
	void TranslateX(float value);
	void TranslateY(float value);
	void TranslateZ(float value);
	void Pitch(float degrees);
	void Yaw(float degrees);

	const FPoint3& GetPosition() const { return m_Position; }
	const FVector3& GetRight() const { return m_Right; }
	const FVector3& GetUp() const { return m_Up; }
	const FVector3& GetForward() const { return m_Forward; }
	const FMatrix4& GetLookAtMatrix() const { return m_LookAtMatrix; }
	const FMatrix4& GetProjectionMatrix() const { return m_ProjectionMatrix; }
	FMatrix4 GetViewMatrix() const { return Inverse(GetLookAtMatrix()); }
	FMatrix3 GetRotationMatrix() const;

	float SetFov(float degrees);
	float GetFov() const { return m_FOV; }
	void GetNearFarValues(float& nearValue, float& farValue);
	void SetNearFarValues(float nearValue, float farValue);
	void SetAspectRatio(float width, float height);
	float GetAspectRatio() const { return m_AspectRatio; }
	void SetMoveSpeed(float value);
	float GetMoveSpeed() const { return m_MoveSpeed; }
	void SetRotationSpeed(float value);
	float GetRotationSpeed() const { return m_RotationSpeed; }

protected:
	float m_FOV, m_AspectRatio, m_Near, m_Far, m_MoveSpeed, m_RotationSpeed;
	FMatrix4 m_ProjectionMatrix, m_LookAtMatrix;
	// views into the columns of the look-at matrix, like the reference's reference members (Camera.cpp:13-16)
	FPoint3& m_Position;
	FVector3& m_Right;
	FVector3& m_Up;
	FVector3& m_Forward;
	void RecalculateProjectionMatrix();
	void RecalculateVectors();
};

// ---- Mesh.h
class Mesh
{
public:
	enum TextureID { Diffuse, Normal, Specular, Glossiness };
	Mesh(const std::vector<IVertex>& vertexBuffer, short stride, short type,
		const std::vector<unsigned int>& indexBuffer, PrimitiveTopology topology, const FPoint3& pos = FPoint3{ 0.f, 0.f, 0.f });
	// takes ownership of both arrays (delete[] in the destructor, Mesh.cpp:32-36)
	Mesh(IVertex* pVertexBuffer, unsigned int numVertices, short stride, short type, unsigned int* pIndexBuffer,
		unsigned int numIndices, PrimitiveTopology topology, const FPoint3& pos = FPoint3{ 0.f, 0.f, 0.f });
	virtual ~Mesh();
	Mesh(const Mesh&) = delete;
	Mesh& operator=(const Mesh&) = delete;

	FPoint3& GetPosition() { return m_Position; }
	FMatrix4& GetWorld() { return m_WorldSpace; }
	const FPoint3& GetPositionConst() const { return m_Position; }
	const FMatrix4& GetWorldConst() const { return m_WorldSpace; }
	FMatrix3 GetRotationMatrix() const { return (FMatrix3)m_WorldSpace; }
	short GetVertexType() const { return m_VertexType; }
	short GetVertexStride() const { return m_VertexStride; }
	PrimitiveTopology GetTopology() const { return m_Topology; }
	unsigned int GetNumVertices() const { return m_NumVertices; }
	unsigned int GetNumIndices() const { return m_NumIndices; }
	unsigned int GetTotalNumTriangles() const;
	IVertex* GetVertexBuffer() const { return m_pVertexBuffer; }
	unsigned int* GetIndexBuffer() const { return m_pIndexBuffer; }
	const int* GetTextureIds() const { return m_TextureIds; }
	void SetTextureIds(int diff, int norm = -1, int spec = -1, int gloss = -1);
	void SetTextureId(int id, TextureID texID = TextureID::Diffuse);

protected:
	short m_VertexType, m_VertexStride;
	PrimitiveTopology m_Topology;
	unsigned int m_NumVertices, m_NumIndices;
	IVertex* m_pVertexBuffer;
	unsigned int* m_pIndexBuffer;
	int m_TextureIds[4];
	FMatrix4 m_WorldSpace;
	FPoint3& m_Position;
};

// ---- Light.h / DirectionalLight.h / PointLight.h (scene data only: the GPU path hard-codes its
// light, CUDAROPs.cu:418-421; lights matter to the ERenderer CPU path)
class Light
{
public:
	explicit Light(const RGBColor& colour, float intensity) : m_Intensity(intensity), m_Colour(colour) {}
	virtual ~Light() = default;
	virtual RGBColor GetBiradiance(const FPoint3& pointToShade) const = 0;
	virtual FVector3 GetDirection(const FPoint3& pointToShade) const = 0;
	virtual float GetIntensity() const { return m_Intensity; }
protected:
	float m_Intensity;
	RGBColor m_Colour;
};
class DirectionalLight : public Light
{
public:
	explicit DirectionalLight(const RGBColor& colour, float intensity, const FVector3& direction) : Light(colour, intensity), m_Direction(direction) {}
	RGBColor GetBiradiance(const FPoint3&) const override { return { m_Colour.r * m_Intensity, m_Colour.g * m_Intensity, m_Colour.b * m_Intensity }; }
	FVector3 GetDirection(const FPoint3&) const override { return m_Direction; }
private:
	FVector3 m_Direction;
};

// ---- SceneGraph.h
class SceneGraph
{
public:
	SceneGraph() = default;
	virtual ~SceneGraph();
	SceneGraph(const SceneGraph&) = delete;
	SceneGraph& operator=(const SceneGraph&) = delete;
	void AddMesh(Mesh* pMesh);       // takes ownership
	void AddLight(Light* pLight);    // takes ownership
	const std::vector<Mesh*>& GetMeshes() const { return m_pMeshes; }
	const std::vector<Light*>& GetLights() const { return m_pLights; }
	void Update(float elapsedSec);
protected:
	std::vector<Mesh*> m_pMeshes;
	std::vector<Light*> m_pLights;
};

// ---- SceneManager.h (the key-driven Update of the reference needs SDL; the toggles are public)
class SceneManager final
{
public:
	SceneManager();
	~SceneManager();
	SceneManager(const SceneManager&) = delete;
	SceneManager& operator=(const SceneManager&) = delete;
	void Update(float elapsedSec);
	void AddSceneGraph(SceneGraph* pSceneGraph);   // takes ownership
	VisualisationState GetVisualisationState() const { return m_Visualisation; }
	SceneGraph* GetSceneGraph() const;
	SampleState GetSampleState() const { return m_SampleState; }
	CullingMode GetCullingMode() const { return m_CullingMode; }
	void ChangeSceneGraph();
	void ChangeSceneGraph(int idx);
	void ToggleVisualisation();
	void ToggleSampleState();
	void ToggleCullingMode();
private:
	int m_Index;
	VisualisationState m_Visualisation;
	std::vector<SceneGraph*> m_pSceneGraphs;
	SampleState m_SampleState;
	CullingMode m_CullingMode;
};

// ---- ObjParser.h : same output buffers as the reference parser (vertex order, indices, tangents),
// with the O(n^2) duplicate search replaced by a hash on (position index, uv index)
class ObjParser
{
public:
	ObjParser();
	virtual ~ObjParser();
	bool OpenFile(const std::string& filePath);
	void CloseFile();
	void ReadFromObjFile(IVertex*& pVertexBuffer, unsigned int& numVertices, unsigned int*& pIndexBuffer, unsigned int& numIndices, short& vertexType);
	void ReadFromObjFile(std::vector<IVertex>& vertexBuffer, std::vector<unsigned int>& indexBuffer, short& vertexType);
	void SetInvertYAxis(bool value);
	void SetVerbose(bool value) { m_Verbose = value; }
private:
	bool m_IsYAxisInverted;
	bool m_Verbose;
	std::ifstream m_ReadFile;
	std::vector<FPoint3> m_Positions;
	std::vector<FVector2> m_UVs;
	std::vector<FVector3> m_Normals;
	std::vector<unsigned int> m_PositionIndices, m_UVIndices, m_NormalIndices;
	unsigned int ParseData();
	void Assign(unsigned int faceCount, std::vector<IVertex>& vertices, std::vector<unsigned int>& indices, short& vertexType, bool pointerFlavour);
};

// ---- CUDATexture.h / CUDATextureManager.h
class CUDATextureManager;
class CUDATexture
{
public:
	CUDATexture();
	CUDATexture(const std::string& path);
	virtual ~CUDATexture() noexcept;
	CUDATexture(const CUDATexture&) = delete;
	CUDATexture& operator=(const CUDATexture&) = delete;
	bool IsAllocated() const { return m_Width != 0 && m_Height != 0; }
	// decodes an 8-bit PNG (the reference goes through SDL_image's IMG_Load, CUDATexture.cpp:28)
	virtual bool Create(const char* path);
	// headless variant: RGBA8 pixels, byte 0 = R
	virtual bool Create(const unsigned int* rgba, unsigned int width, unsigned int height);
	virtual void Destroy();
	unsigned int GetBytesPerPixel() const { return m_BytesPerPixel; }
	unsigned int GetWidth() const { return m_Width; }
	unsigned int GetHeight() const { return m_Height; }
	const std::vector<unsigned int>& GetPixels() const { return m_Pixels; }
protected:
	unsigned int m_BytesPerPixel, m_Width, m_Height;
	std::vector<unsigned int> m_Pixels;   // host copy; uploaded to a renderer at LoadScene
};

class CUDATextureManager final
{
public:
	CUDATextureManager();
	~CUDATextureManager();
	CUDATextureManager(const CUDATextureManager&) = delete;
	CUDATextureManager& operator=(const CUDATextureManager&) = delete;
	CUDATexture* GetCUDATexture(int id) const;
	int AddCUDATexture(CUDATexture* pTex);    // takes ownership; throws std::runtime_error on null
	void RemoveCUDATexture(int id);
	unsigned int GetNumTextures() const { return m_CurrentTextureId; }
private:
	unsigned int m_CurrentTextureId;
	std::vector<CUDATexture*> m_pCUDATextures;
};

// ---- CUDABenchMarker.h : StartTimer / StopTimer around a frame (the reference times with a cudaEvent pair and averages;
// here the renderer's own per-stage device timers are read, StopTimer returns the last frame's total in ms)
class CUDARenderer;
class CUDABenchMarker
{
public:
	explicit CUDABenchMarker(int amountOfTimes = 4, int amountOfTimeAverages = 10) : m_AmountOfTimes(amountOfTimes), m_AmountOfTimeAverages(amountOfTimeAverages) {}
	virtual ~CUDABenchMarker() = default;
	virtual void StartTimer();
	virtual float StopTimer();
private:
	friend class CUDARenderer;
	int m_AmountOfTimes, m_AmountOfTimeAverages;
	CUDARenderer* m_pRenderer{};
};

// ---- CUDARenderer.h
class CUDARenderer final
{
public:
	CUDARenderer(const WindowHelper& windowHelper);
	~CUDARenderer() noexcept;
	CUDARenderer(const CUDARenderer&) = delete;
	CUDARenderer& operator=(const CUDARenderer&) = delete;

	// testing purposes (CUDARenderer.h:43-45): the texture file drawn over the window, then Present
	void DrawTexture(char* tp);
	void DrawTextureGlobal(char* tp, bool isStretchedToWindow = true, SampleState sampleState = (SampleState)0);

	void LoadScene(const SceneGraph* pSceneGraph, const CUDATextureManager& tm);
	void SetupBins(const IPoint2& numBins, const IPoint2& binDim, unsigned int binQueueMaxSize);
	int EnterValidRenderingState();
	void Render(const SceneManager& sm, const CUDATextureManager& tm, const Camera* pCamera);
	void Present();
	void RenderAuto(const SceneManager& sm, const CUDATextureManager& tm, const Camera* pCamera);
	unsigned int GetTotalNumVisibleTriangles() const;
	unsigned int GetTotalNumTriangles() const { return m_TotalNumTriangles; }
	void DisplayGPUSpecs(int deviceId = 0);
	void KernelWarmUp();
	void StartTimer();
	float StopTimer();
	CUDABenchMarker& GetBenchMarker() { return m_BenchMarker; }

	// additions of the headless build
	// The reference snapshots every mesh's world matrix when the scene is loaded (CUDAMesh.cpp:16-18), so its GPU path never
	// sees SceneGraph::Update.  Here Render() compares the meshes' current world matrices with the ones the device holds
	// and refreshes those that changed (b2r_mesh_set_world; the per-mesh WVP upload of CUDARenderer.cu:241-246 is then
	// rebuilt on the device).  With unchanged matrices - the reference's main passes elapsedSec = 0 - nothing differs.
	void SetDynamicWorlds(bool enabled) { m_DynamicWorlds = enabled; }
	b2r_context* GetContext() const { return m_pContext; }
	void SetClearColour(const RGBColor& colour) { m_ClearColour = colour; }
	const char* GetLastError() const;

private:
	unsigned int m_TotalNumTriangles{};
	const WindowHelper& m_WindowHelper;
	b2r_context* m_pContext{};
	RGBColor m_ClearColour{ 0.25f, 0.25f, 0.25f };
	CUDABenchMarker m_BenchMarker{};
	const SceneGraph* m_pSceneGraph{};              // not owned: the graph of the last LoadScene, for the world-matrix refresh
	std::vector<FMatrix4> m_DeviceWorlds;           // world matrix the device holds for mesh i
	std::vector<int> m_MeshIds, m_ContextTextureIds;
	bool m_DynamicWorlds{ true };
	void DrawTextureImpl(const char* tp, int global, bool stretched, SampleState sampleState);
	void* m_TimerStart{};
	void* m_TimerStop{};
};

// ---- ERenderer.h : the reference's CPU software renderer entry point, kept as an adapter that
// drives the same B200 path (this build has no CPU rendering path)
namespace Elite
{
	class Renderer final
	{
	public:
		// ERenderer.h:28.  This build has no SDL to ask the window for its size or surface: the renderer then owns a back
		// buffer of DefaultWidth x DefaultHeight (640 x 480, Resolution::SD, the reference's default; main.cpp:405) and the
		// window pointer is only kept.  GetBackBufferPixels() is the frame.
		Renderer(SDL_Window* pWindow);
		Renderer(const WindowHelper& windowHelper);
		static unsigned int DefaultWidth, DefaultHeight;
		const unsigned int* GetBackBufferPixels() const { return m_WindowHelper.pBackBufferPixels; }
		~Renderer();
		Renderer(const Renderer&) = delete;
		Renderer& operator=(const Renderer&) = delete;
		void LoadScene(const SceneGraph* pSceneGraph, const CUDATextureManager& tm);
		void Render(const SceneManager& sm);
		void SetCamera(Camera* pCamera) { m_pCamera = pCamera; }
		bool SaveBackbufferToImage() const;     // writes BackbufferRender.bmp like ERenderer.cpp:174-177
	private:
		std::vector<unsigned int> m_OwnPixels;      // SDL_Window* constructor only
		WindowHelper m_OwnHelper;
		const WindowHelper& m_WindowHelper;
		CUDARenderer m_Renderer;
		const CUDATextureManager* m_pTextures{};
		Camera* m_pCamera{};
	};
}
