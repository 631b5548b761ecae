// Implementation of the C++ veneer (b2r_scene.hpp) on top of the C ABI (include/b2r.h).
// Each method cites the reference method whose observable behaviour it keeps
// (paths relative to /root/reference/CustomCUDARasterizer).
#include "b2r_scene.hpp"
#include "../../../include/b2r.h"

#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <memory>
#include <unordered_map>

// =========================================================================== Resolution
Resolution Resolution::GetResolution(ResolutionStandard standard)   // Resolution.cpp:4-18
{
	static const UInt2 sizes[] = { { 320U, 240U }, { 640U, 480U }, { 1280U, 720U }, { 1920U, 1080U }, { 2560U, 1440U }, { 3840U, 2160U } };
	static const UInt2 aspects[] = { { 4, 3 }, { 4, 3 }, { 16, 9 }, { 16, 9 }, { 16, 9 }, { 16, 9 } };
	Resolution res;
	res.Width = sizes[standard].w; res.Height = sizes[standard].h;
	res.AspectRatio = aspects[standard]; res.Standard = standard;
	return res;
}

// =========================================================================== Camera
Camera::Camera(const FPoint3& position, float fov)   // Camera.cpp:6-25
	: m_FOV{ 0.f }, m_AspectRatio{}, m_Near{ 0.1f }, m_Far{ 100.f }, m_MoveSpeed{ 5.f }, m_RotationSpeed{ 1.f }
	, m_ProjectionMatrix{ FMatrix4::Identity() }, m_LookAtMatrix{ FMatrix4::Identity() }
	, m_Position{ reinterpret_cast<FPoint3&>(m_LookAtMatrix.data[3][0]) }
	, m_Right{ reinterpret_cast<FVector3&>(m_LookAtMatrix.data[0][0]) }
	, m_Up{ reinterpret_cast<FVector3&>(m_LookAtMatrix.data[1][0]) }
	, m_Forward{ reinterpret_cast<FVector3&>(m_LookAtMatrix.data[2][0]) }
{
	SetFov(fov);
	m_Position = position;
}

FMatrix3 Camera::GetRotationMatrix() const
{
	FMatrix3 m;
	m.data[0][0] = m_Right.x; m.data[0][1] = m_Right.y; m.data[0][2] = m_Right.z;
	m.data[1][0] = m_Up.x; m.data[1][1] = m_Up.y; m.data[1][2] = m_Up.z;
	m.data[2][0] = m_Forward.x; m.data[2][1] = m_Forward.y; m.data[2][2] = m_Forward.z;
	return m;
}

float Camera::SetFov(float degrees)   // Camera.cpp:27-32: stores tan(fov/2)
{
	const float radians = ToRadians(degrees);
	m_FOV = tanf(radians / 2.f);
	return m_FOV;
}
void Camera::SetAspectRatio(float width, float height) { m_AspectRatio = width / height; RecalculateProjectionMatrix(); }
void Camera::GetNearFarValues(float& nearValue, float& farValue) { nearValue = m_Near; farValue = m_Far; }
void Camera::SetNearFarValues(float nearValue, float farValue) { m_Near = nearValue; m_Far = farValue; RecalculateProjectionMatrix(); }
void Camera::TranslateX(float value) { m_Position += m_Right * value; }
void Camera::TranslateY(float value) { m_Position += FVector3{ 0.f, 1.f, 0.f } * value; }
void Camera::TranslateZ(float value) { m_Position += m_Forward * value; }
void Camera::Pitch(float degrees)   // Camera.cpp:67-73
{
	const FMatrix3 rotX = MakeRotationX(ToRadians(degrees));
	m_Forward = rotX * m_Forward;
	RecalculateVectors();
}
void Camera::Yaw(float degrees)     // Camera.cpp:75-81
{
	const FMatrix3 rotY = MakeRotationY(ToRadians(degrees));
	m_Forward = rotY * m_Forward;
	RecalculateVectors();
}
static float ClampF(float a, float lo, float hi) { if (a < lo) return lo; if (a > hi) return hi; return a; }
void Camera::SetMoveSpeed(float value) { m_MoveSpeed = ClampF(m_MoveSpeed + value, 1.f, 25.f); }
void Camera::SetRotationSpeed(float value) { m_RotationSpeed = ClampF(m_RotationSpeed + value, 0.1f, 5.f); }
void Camera::RecalculateProjectionMatrix()   // Camera.cpp:97-108 (m(3,3) stays 1: w_clip = -z_view + 1)
{
	const float A = m_Far / (m_Near - m_Far);
	const float B = (m_Far * m_Near) / (m_Near - m_Far);
	const float value = 1.f / (m_AspectRatio * m_FOV);
	const float invFov = 1.f / m_FOV;
	m_ProjectionMatrix.data[0][0] = value;
	m_ProjectionMatrix.data[1][1] = invFov;
	m_ProjectionMatrix.data[2][2] = A;
	m_ProjectionMatrix.data[2][3] = -1.f;
	m_ProjectionMatrix.data[3][2] = B;
}
void Camera::RecalculateVectors()   // Camera.cpp:110-117
{
	Normalize(m_Forward);
	m_Right = Cross(FVector3{ 0.f, 1.f, 0.f }, m_Forward);
	Normalize(m_Right);
	m_Up = Cross(m_Forward, m_Right);
	Normalize(m_Up);
}

// =========================================================================== Mesh
Mesh::Mesh(IVertex* pVertexBuffer, unsigned int numVertices, short stride, short type, unsigned int* pIndexBuffer,
	unsigned int numIndices, PrimitiveTopology topology, const FPoint3& pos)   // Mesh.cpp:15-30
	: m_VertexType{ type }, m_VertexStride{ stride }, m_Topology{ topology }, m_NumVertices{ numVertices }, m_NumIndices{ numIndices }
	, m_pVertexBuffer{ pVertexBuffer }, m_pIndexBuffer{ pIndexBuffer }, m_TextureIds{ -1, -1, -1, -1 }
	, m_WorldSpace{ FMatrix4::Identity() }, m_Position{ reinterpret_cast<FPoint3&>(m_WorldSpace.data[3][0]) }
{
	m_Position += FVector3{ pos.x, pos.y, pos.z };
}

Mesh::Mesh(const std::vector<IVertex>& vertexBuffer, short stride, short type,
	const std::vector<unsigned int>& indexBuffer, PrimitiveTopology topology, const FPoint3& pos)   // Mesh.cpp:4-13
	: Mesh{ nullptr, (unsigned int)vertexBuffer.size(), stride, type, nullptr, (unsigned int)indexBuffer.size(), topology, pos }
{
	m_pVertexBuffer = new IVertex[m_NumVertices];
	if (m_NumVertices) memcpy((void*)m_pVertexBuffer, vertexBuffer.data(), sizeof(IVertex) * m_NumVertices);     // (empty vector: data() may be null)
	m_pIndexBuffer = new unsigned int[m_NumIndices];
	if (m_NumIndices) memcpy(m_pIndexBuffer, indexBuffer.data(), sizeof(unsigned int) * m_NumIndices);
	// the reference's copying constructor translates a second time after delegating (Mesh.cpp:12)
	m_Position += FVector3{ pos.x, pos.y, pos.z };
}

Mesh::~Mesh() { delete[] m_pVertexBuffer; delete[] m_pIndexBuffer; }
void Mesh::SetTextureIds(int diff, int norm, int spec, int gloss) { m_TextureIds[0] = diff; m_TextureIds[1] = norm; m_TextureIds[2] = spec; m_TextureIds[3] = gloss; }
void Mesh::SetTextureId(int id, TextureID texID) { m_TextureIds[texID] = id; }
unsigned int Mesh::GetTotalNumTriangles() const   // Mesh.cpp:51-64
{
	switch (m_Topology)
	{
	case PrimitiveTopology::TriangleList: return m_NumIndices / 3;
	case PrimitiveTopology::TriangleStrip: return m_NumIndices - 2;
	}
	return 0;
}

// =========================================================================== SceneGraph / SceneManager
SceneGraph::~SceneGraph()   // SceneGraph.cpp:6-18
{
	for (Mesh* m : m_pMeshes) delete m;
	for (Light* l : m_pLights) delete l;
}
void SceneGraph::AddMesh(Mesh* pMesh) { m_pMeshes.push_back(pMesh); }
void SceneGraph::AddLight(Light* pLight) { m_pLights.push_back(pLight); }
void SceneGraph::Update(float elapsedSec)   // SceneGraph.cpp:30-39
{
	const FMatrix4 rotationMatrix = MakeRotationY4(1.f * elapsedSec);
	for (Mesh* m : m_pMeshes) m->GetWorld() *= rotationMatrix;
}

SceneManager::SceneManager()   // SceneManager.cpp:11-17
	: m_Index{ 0 }, m_Visualisation{ VisualisationState::PBR }, m_pSceneGraphs{}, m_SampleState{ SampleState::Point }, m_CullingMode{ CullingMode::BackFace } {}
SceneManager::~SceneManager() { for (SceneGraph* g : m_pSceneGraphs) delete g; }
void SceneManager::AddSceneGraph(SceneGraph* pSceneGraph) { m_pSceneGraphs.push_back(pSceneGraph); }
SceneGraph* SceneManager::GetSceneGraph() const   // SceneManager.cpp:33-44
{
	const size_t size = m_pSceneGraphs.size();
	if (size == 0)
	{
		std::cout << "\n!Warning; no SceneGraph(s) detected in the SceneManager!\n";
		return nullptr;
	}
	if ((size_t)m_Index >= size) return m_pSceneGraphs.at(size - 1);
	return m_pSceneGraphs.at(m_Index);
}
void SceneManager::ChangeSceneGraph() { ++m_Index; if ((size_t)m_Index >= m_pSceneGraphs.size()) m_Index = 0; }
void SceneManager::ChangeSceneGraph(int idx) { if (idx > 0 && (size_t)idx < m_pSceneGraphs.size()) m_Index = idx; }
void SceneManager::Update(float elapsedSec)   // SceneManager.cpp:60-87 minus the SDL key polling
{
	for (SceneGraph* g : m_pSceneGraphs) g->Update(elapsedSec);
}
void SceneManager::ToggleVisualisation()
{
	m_Visualisation = VisualisationState((int)m_Visualisation + 1);
	if ((int)m_Visualisation > 2) m_Visualisation = VisualisationState::PBR;
}
void SceneManager::ToggleSampleState()
{
	m_SampleState = SampleState((int)m_SampleState + 1);
	if ((int)m_SampleState > 1) m_SampleState = SampleState::Point;
}
void SceneManager::ToggleCullingMode()
{
	m_CullingMode = CullingMode((int)m_CullingMode + 1);
	if ((int)m_CullingMode > 2) m_CullingMode = CullingMode::BackFace;
}

// =========================================================================== ObjParser
ObjParser::ObjParser() : m_IsYAxisInverted{}, m_Verbose{} {}
ObjParser::~ObjParser() { CloseFile(); }
void ObjParser::SetInvertYAxis(bool value) { m_IsYAxisInverted = value; }
void ObjParser::CloseFile() { m_ReadFile.close(); }

bool ObjParser::OpenFile(const std::string& filePath)   // ObjParser.cpp:564-584
{
	const size_t dot = filePath.find('.');
	if (dot == std::string::npos) return false;
	if (filePath.substr(dot) != ".obj")
	{
		std::cout << "Only '.obj' file formats are supported\n Current file extension is : " << filePath.substr(dot) << '\n';
		return false;
	}
	if (m_ReadFile.is_open()) CloseFile();
	m_ReadFile.clear();
	m_ReadFile.open(filePath);
	return m_ReadFile.is_open();
}

namespace {
// "a b c" -> three space-delimited tokens, like ObjParser::GetFirstSecondThird (ObjParser.cpp:196-208)
void SplitThree(const std::string& s, std::string& a, std::string& b, std::string& c)
{
	std::stringstream ss{ s };
	std::getline(ss, a, ' '); std::getline(ss, b, ' '); std::getline(ss, c, ' ');
}
}

unsigned int ObjParser::ParseData()   // ObjParser.cpp:23-87
{
	m_Positions.clear(); m_UVs.clear(); m_Normals.clear();
	m_PositionIndices.clear(); m_UVIndices.clear(); m_NormalIndices.clear();
	if (!m_ReadFile.is_open()) return 0;
	unsigned int faceCount = 0;
	std::string line, a, b, c;
	std::vector<unsigned int> faceIndexes;
	while (std::getline(m_ReadFile, line))
	{
		if (line.empty() || line.front() == '#') continue;
		const std::string prefix = line.substr(0, line.find(' '));
		if (prefix == "v")
		{
			// the reference cuts three characters: 'v' and TWO spaces (3ds-Max style), ObjParser.cpp:45
			SplitThree(line.substr(3), a, b, c);
			m_Positions.push_back(FPoint3{ std::stof(a), std::stof(b), std::stof(c) });
		}
		else if (prefix == "f")
		{
			SplitThree(line.substr(2), a, b, c);
			const std::string* corner[3] = { &a, &b, &c };
			faceIndexes.clear();
			for (int i = 0; i < 3; ++i)
			{
				// indices of one corner, '/'-separated, OBJ is 1-based (ObjParser.cpp:131-146)
				size_t begin = 0;
				for (;;)
				{
					const size_t slash = corner[i]->find('/', begin);
					faceIndexes.push_back((unsigned int)(std::stoi(corner[i]->substr(begin, slash == std::string::npos ? std::string::npos : slash - begin)) - 1));
					if (slash == std::string::npos) break;
					begin = slash + 1;
				}
			}
			const size_t size = faceIndexes.size(), per = size / 3;
			for (size_t idx = 0; idx < size; ++idx)
				switch (idx % per)
				{
				case 0: m_PositionIndices.push_back(faceIndexes[idx]); break;
				case 1: m_UVIndices.push_back(faceIndexes[idx]); break;
				case 2: m_NormalIndices.push_back(faceIndexes[idx]); break;
				}
			++faceCount;
		}
		else if (prefix == "vt")
		{
			SplitThree(line.substr(3), a, b, c);
			float y = std::stof(b);
			if (m_IsYAxisInverted) y = 1 - y;
			m_UVs.push_back(FVector2{ std::stof(a), y });
		}
		else if (prefix == "vn")
		{
			SplitThree(line.substr(3), a, b, c);
			m_Normals.push_back(GetNormalized(FVector3{ std::stof(a), std::stof(b), std::stof(c) }));
		}
		else if (m_Verbose)
			std::cout << "ObjParser: skipped line '" << line << "'\n";
	}
	return faceCount;
}

// ObjParser.cpp:210-379 / 381-548.  Vertices are numbered by first occurrence; two corners share a
// vertex when their position index and uv index are equal and their normals compare equal under
// the reference's tolerant Vector== (Vector3.h:106-107).  Instead of scanning every earlier vertex
// the candidates are looked up by (position index, uv index) and visited in creation order, which
// finds the same first match.
void ObjParser::Assign(unsigned int faceCount, std::vector<IVertex>& vertices, std::vector<unsigned int>& indices, short& vertexType, bool pointerFlavour)
{
	const size_t numIndices = m_PositionIndices.size();
	vertices.clear(); indices.clear();
	indices.reserve(numIndices);
	const bool isUVs = !m_UVs.empty(), isNormals = !m_Normals.empty();
	vertexType |= isUVs * (int)VertexType::Uv;
	vertexType |= isNormals * (int)VertexType::Norm;

	struct Entry { unsigned int normIdx, vertexIdx; };
	std::unordered_map<uint64_t, std::vector<Entry>> seen;
	seen.reserve(numIndices);
	auto key = [&](size_t i) { return ((uint64_t)m_PositionIndices[i] << 32) | (isUVs ? m_UVIndices[i] : 0u); };
	auto normalsEqual = [&](unsigned int a, unsigned int b)
	{
		const FVector3& x = m_Normals[a]; const FVector3& y = m_Normals[b];
		return AreEqual(x.x, y.x) && AreEqual(x.y, y.y) && AreEqual(x.z, y.z);
	};

	if (!(isUVs && isNormals) && !(isUVs && !isNormals) && !(!isUVs && !isNormals))
		return;   // normals without uvs: the reference has no branch for it and leaves the buffers empty
	// file-supplied indices (1-based in the file; index 0 wraps to 0xFFFFFFFF): the reference reads out of bounds here,
	// a library must not - a face that refers to data the file does not hold leaves the buffers empty
	for (size_t i = 0; i < numIndices; ++i)
		if (m_PositionIndices[i] >= m_Positions.size() || (isUVs && (i >= m_UVIndices.size() || m_UVIndices[i] >= m_UVs.size())) ||
			(isUVs && isNormals && (i >= m_NormalIndices.size() || m_NormalIndices[i] >= m_Normals.size())))
			return;
	if (isUVs && isNormals) vertexType |= (int)VertexType::Tan;

	for (size_t i = 0; i < numIndices; ++i)
	{
		std::vector<Entry>& bucket = seen[key(i)];
		const Entry* hit = nullptr;
		for (const Entry& e : bucket)
			if (!(isUVs && isNormals) || normalsEqual(e.normIdx, m_NormalIndices[i])) { hit = &e; break; }
		if (hit) { indices.push_back(hit->vertexIdx); continue; }
		IVertex v;
		v.p = m_Positions[m_PositionIndices[i]];
		if (isUVs)
		{
			v.uv = m_UVs[m_UVIndices[i]];
			if (isNormals) v.n = m_Normals[m_NormalIndices[i]];
		}
		else
			v.c = RGBColor{ 1.f, 1.f, 1.f };
		bucket.push_back(Entry{ isNormals ? m_NormalIndices[i] : 0u, (unsigned int)vertices.size() });
		indices.push_back((unsigned int)vertices.size());
		vertices.push_back(v);
	}

	if (isUVs && isNormals)
	{
		// per-triangle tangents accumulated in triangle order, then Gram-Schmidt against the normal
		for (size_t i = 0; i + 2 < numIndices; i += 3)
		{
			IVertex& v0 = vertices[indices[i]]; IVertex& v1 = vertices[indices[i + 1]]; IVertex& v2 = vertices[indices[i + 2]];
			const FVector3 edge0 = v1.p - v0.p, edge1 = v2.p - v0.p;
			const FVector2 diffX{ v1.uv.x - v0.uv.x, v2.uv.x - v0.uv.x };
			const FVector2 diffY{ v1.uv.y - v0.uv.y, v2.uv.y - v0.uv.y };
			const float r = 1.f / Cross(diffX, diffY);
			const FVector3 tangent = (edge0 * diffY.y - edge1 * diffY.x) * r;
			v0.tan += tangent; v1.tan += tangent; v2.tan += tangent;
		}
		for (IVertex& v : vertices) v.tan = GetNormalized(Reject(v.tan, v.n));
	}
	else if (pointerFlavour && vertices.size() < faceCount)
	{
		// the raw-pointer overload sizes the buffer by the face count and only shrinks it on the
		// textured path (ObjParser.cpp:212,269-277 vs :308-378): the tail stays zero-initialised
		vertices.resize(faceCount);
		for (size_t i = 0; i < vertices.size(); ++i) { /* IVertex() is all zeros, like the reference's new IVertex[n] */ }
	}
}

void ObjParser::ReadFromObjFile(std::vector<IVertex>& vertexBuffer, std::vector<unsigned int>& indexBuffer, short& vertexType)
{
	const unsigned int faceCount = ParseData();
	Assign(faceCount, vertexBuffer, indexBuffer, vertexType, false);
}

void ObjParser::ReadFromObjFile(IVertex*& pVertexBuffer, unsigned int& numVertices, unsigned int*& pIndexBuffer, unsigned int& numIndices, short& vertexType)
{
	const unsigned int faceCount = ParseData();
	std::vector<IVertex> v; std::vector<unsigned int> idx;
	Assign(faceCount, v, idx, vertexType, true);
	numVertices = (unsigned int)v.size(); numIndices = (unsigned int)idx.size();
	pVertexBuffer = new IVertex[numVertices];
	pIndexBuffer = new unsigned int[numIndices];
	if (numVertices) memcpy((void*)pVertexBuffer, v.data(), sizeof(IVertex) * numVertices);
	if (numIndices) memcpy(pIndexBuffer, idx.data(), sizeof(unsigned int) * numIndices);
}

// =========================================================================== PNG (8-bit RGB / RGBA, non-interlaced)
namespace {
uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
bool DecodePNG(const char* path, std::vector<unsigned int>& out, unsigned int& w, unsigned int& h)
{
	FILE* f = fopen(path, "rb");
	if (!f) return false;
	std::vector<unsigned char> file;
	unsigned char buf[65536]; size_t n;
	while ((n = fread(buf, 1, sizeof buf, f)) > 0) file.insert(file.end(), buf, buf + n);
	fclose(f);
	static const unsigned char sig[8] = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };
	if (file.size() < 33 || memcmp(file.data(), sig, 8) != 0) return false;
	std::vector<unsigned char> idat;
	unsigned int bitDepth = 0, colourType = 0, interlace = 0;
	for (size_t off = 8; off + 12 <= file.size();)
	{
		const uint32_t len = be32(&file[off]);
		const unsigned char* type = &file[off + 4];
		if (off + 12 + len > file.size()) return false;
		if (!memcmp(type, "IHDR", 4) && len < 13) return false;
		if (!memcmp(type, "IHDR", 4)) { w = be32(&file[off + 8]); h = be32(&file[off + 12]); bitDepth = file[off + 16]; colourType = file[off + 17]; interlace = file[off + 20]; }
		else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), &file[off + 8], &file[off + 8] + len);
		else if (!memcmp(type, "IEND", 4)) break;
		off += 12 + (size_t)len;
	}
	if (bitDepth != 8 || interlace != 0 || (colourType != 6 && colourType != 2) || !w || !h) return false;
	if (w > 32768u || h > 32768u) return false;      // dimensions come from the file: keep the allocations below sane
	const unsigned int bpp = colourType == 6 ? 4 : 3;
	const size_t stride = (size_t)w * bpp;
	std::vector<unsigned char> raw((stride + 1) * h);
	uLongf rawLen = (uLongf)raw.size();
	if (uncompress(raw.data(), &rawLen, idat.data(), (uLong)idat.size()) != Z_OK || rawLen != raw.size()) return false;
	std::vector<unsigned char> img(stride * h);
	for (unsigned int y = 0; y < h; ++y)
	{
		const unsigned char ft = raw[(stride + 1) * y];
		const unsigned char* src = &raw[(stride + 1) * y + 1];
		unsigned char* dst = &img[stride * y];
		const unsigned char* up = y ? &img[stride * (y - 1)] : nullptr;
		for (size_t i = 0; i < stride; ++i)
		{
			const int a = i >= bpp ? dst[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
			int pred = 0;
			switch (ft)
			{
			case 0: pred = 0; break;
			case 1: pred = a; break;
			case 2: pred = b; break;
			case 3: pred = (a + b) >> 1; break;
			case 4: { const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
			default: return false;
			}
			dst[i] = (unsigned char)(src[i] + pred);
		}
	}
	out.resize((size_t)w * h);
	for (size_t p = 0; p < (size_t)w * h; ++p)
	{
		const unsigned char* s = &img[p * bpp];
		out[p] = (uint32_t)s[0] | ((uint32_t)s[1] << 8) | ((uint32_t)s[2] << 16) | ((uint32_t)(bpp == 4 ? s[3] : 255) << 24);
	}
	return true;
}
}

// =========================================================================== CUDATexture / CUDATextureManager
CUDATexture::CUDATexture() : m_BytesPerPixel{}, m_Width{}, m_Height{} {}
CUDATexture::CUDATexture(const std::string& path) : CUDATexture{} { Create(path.c_str()); }
CUDATexture::~CUDATexture() noexcept { Destroy(); }
bool CUDATexture::Create(const char* path)   // CUDATexture.cpp:24-72: false when the image cannot be loaded
{
	Destroy();
	std::vector<unsigned int> px; unsigned int w = 0, h = 0;
	if (!path || !DecodePNG(path, px, w, h)) return false;
	m_Pixels.swap(px); m_Width = w; m_Height = h; m_BytesPerPixel = 4;
	return true;
}
bool CUDATexture::Create(const unsigned int* rgba, unsigned int width, unsigned int height)
{
	Destroy();
	if (!rgba || !width || !height) return false;
	m_Pixels.assign(rgba, rgba + (size_t)width * height);
	m_Width = width; m_Height = height; m_BytesPerPixel = 4;
	return true;
}
void CUDATexture::Destroy() { m_Pixels.clear(); m_Width = m_Height = 0; m_BytesPerPixel = 0; }

CUDATextureManager::CUDATextureManager() : m_CurrentTextureId{}, m_pCUDATextures{} { m_pCUDATextures.reserve(4); }
CUDATextureManager::~CUDATextureManager() { for (CUDATexture* t : m_pCUDATextures) delete t; }
CUDATexture* CUDATextureManager::GetCUDATexture(int id) const   // CUDATextureManager.cpp:21-27
{
	if (id >= 0 && m_CurrentTextureId > (unsigned int)id && (size_t)id < m_pCUDATextures.size()) return m_pCUDATextures[id];
	return nullptr;
}
int CUDATextureManager::AddCUDATexture(CUDATexture* pTex)   // CUDATextureManager.cpp:29-37
{
	if (!pTex) throw std::runtime_error{ "CUDATextureManager::AddCUDATexture > pTex is invalid" };
	m_pCUDATextures.push_back(pTex);
	return (int)m_CurrentTextureId++;
}
void CUDATextureManager::RemoveCUDATexture(int id)   // CUDATextureManager.cpp:39-56
{
	CUDATexture* pTex = GetCUDATexture(id);
	if (!pTex) return;
	const auto it = std::find(m_pCUDATextures.cbegin(), m_pCUDATextures.cend(), pTex);
	if (it == m_pCUDATextures.cend()) return;
	m_pCUDATextures.erase(it);
	delete pTex;
}

// =========================================================================== CUDARenderer
CUDARenderer::CUDARenderer(const WindowHelper& windowHelper)   // CUDARenderer.cu:34-46
	: m_WindowHelper{ windowHelper }
{
	m_BenchMarker.m_pRenderer = this;
	const int r = b2r_create(0, windowHelper.Resolution.Width, windowHelper.Resolution.Height, &m_pContext);
	if (r != B2R_OK)
		std::cout << "!CUDARenderer > " << b2r_last_error(nullptr) << "!\n";   // errors are reported, never thrown (PCH.h:46-56)
}
CUDARenderer::~CUDARenderer() noexcept { b2r_destroy(m_pContext); }
const char* CUDARenderer::GetLastError() const { return b2r_last_error(m_pContext); }

void CUDARenderer::LoadScene(const SceneGraph* pSceneGraph, const CUDATextureManager& tm)   // CUDARenderer.cu:178-197
{
	if (!pSceneGraph)
	{
		std::cout << "!CUDARenderer::LoadScene > Invalid scenegraph!\n";
		return;
	}
	if (!m_pContext) return;
	m_TotalNumTriangles = 0;
	b2r_mesh_clear_all(m_pContext);
	// the device textures of the previous LoadScene go first: in the reference they belong to the texture manager and are
	// not re-created per scene; here every LoadScene uploads the manager's current textures
	for (int id : m_ContextTextureIds) b2r_texture_destroy(m_pContext, id);
	m_ContextTextureIds.clear();
	m_pSceneGraph = pSceneGraph; m_DeviceWorlds.clear(); m_MeshIds.clear();
	// textures: the manager's dense ids are kept (CUDATextureManager.cpp:29-37)
	std::vector<int> ctxId(tm.GetNumTextures(), -1);
	for (unsigned int id = 0; id < tm.GetNumTextures(); ++id)
	{
		const CUDATexture* t = tm.GetCUDATexture((int)id);
		if (t && t->IsAllocated() && b2r_texture_create_rgba8(m_pContext, t->GetPixels().data(), t->GetWidth(), t->GetHeight(), &ctxId[id]) == B2R_OK)
			m_ContextTextureIds.push_back(ctxId[id]);
	}
	for (const Mesh* pMesh : pSceneGraph->GetMeshes())
	{
		// CUDAMesh copies the mesh's world matrix and then adds the mesh position to its translation
		// once more (CUDAMesh.cpp:16-18): the device copy is translated twice.
		FMatrix4 world = pMesh->GetWorldConst();
		const FPoint3& pos = pMesh->GetPositionConst();
		world.data[3][0] += pos.x; world.data[3][1] += pos.y; world.data[3][2] += pos.z;
		int tex[4];
		for (int i = 0; i < 4; ++i)
		{
			const int id = pMesh->GetTextureIds()[i];
			tex[i] = (id >= 0 && (size_t)id < ctxId.size()) ? ctxId[id] : -1;
		}
		int meshId = -1;
		const int r = b2r_mesh_upload(m_pContext, pMesh->GetVertexBuffer(), pMesh->GetNumVertices(), pMesh->GetIndexBuffer(),
			pMesh->GetNumIndices(), (int)pMesh->GetTopology(), &world.data[0][0], tex, &meshId);
		m_DeviceWorlds.push_back(world); m_MeshIds.push_back(r == B2R_OK ? meshId : -1);
		if (r != B2R_OK) { std::cout << "!CUDARenderer::LoadScene > " << b2r_last_error(m_pContext) << "!\n"; continue; }
		m_TotalNumTriangles += pMesh->GetTotalNumTriangles();
	}
}

void CUDARenderer::SetupBins(const IPoint2& numBins, const IPoint2& binDim, unsigned int binQueueMaxSize)   // CUDARenderer.cu:200-204
{
	if (!m_pContext) return;
	if (b2r_setup_bins(m_pContext, (uint32_t)numBins.x, (uint32_t)numBins.y, (uint32_t)binDim.x, (uint32_t)binDim.y, binQueueMaxSize) != B2R_OK)
		std::cout << "!CUDARenderer::SetupBins > " << b2r_last_error(m_pContext) << "!\n";
}

int CUDARenderer::EnterValidRenderingState() { return 0; }   // clears are tile-local initialisation inside the frame (CUDARenderer.cu:374-381)

void CUDARenderer::Render(const SceneManager& sm, const CUDATextureManager&, const Camera* pCamera)   // CUDARenderer.cu:207-354
{
	if (!m_pContext || !pCamera) return;
	const FPoint3& camPos = pCamera->GetPosition();
	const FVector3& camFwd = pCamera->GetForward();
	const FMatrix4 viewMatrix = pCamera->GetViewMatrix();
	const FMatrix4 viewProjectionMatrix = pCamera->GetProjectionMatrix() * viewMatrix;
	const float pos[3] = { camPos.x, camPos.y, camPos.z }, fwd[3] = { camFwd.x, camFwd.y, camFwd.z };
	const float clear[3] = { m_ClearColour.r, m_ClearColour.g, m_ClearColour.b };
	if (m_DynamicWorlds && m_pSceneGraph && m_pSceneGraph->GetMeshes().size() == m_DeviceWorlds.size())
	{
		// SceneGraph::Update (SceneGraph.cpp:30-39) / anybody who moved a mesh: the device gets the matrices that changed,
		// with CUDAMesh's translation rule (CUDAMesh.cpp:16-18) applied to the current matrix
		size_t i = 0;
		for (const Mesh* pMesh : m_pSceneGraph->GetMeshes())
		{
			FMatrix4 world = pMesh->GetWorldConst();
			const FPoint3& pos = pMesh->GetPositionConst();
			world.data[3][0] += pos.x; world.data[3][1] += pos.y; world.data[3][2] += pos.z;
			if (m_MeshIds[i] >= 0 && memcmp(&world.data[0][0], &m_DeviceWorlds[i].data[0][0], 64) != 0)
			{
				if (b2r_mesh_set_world(m_pContext, m_MeshIds[i], &world.data[0][0]) == B2R_OK) m_DeviceWorlds[i] = world;
			}
			++i;
		}
	}
	int r = b2r_render(m_pContext, pos, fwd, &viewProjectionMatrix.data[0][0], (int)sm.GetCullingMode(), (int)sm.GetSampleState(),
		(int)sm.GetVisualisationState(), clear);
	if (r == B2R_OK) r = b2r_sync(m_pContext);   // Render is synchronous on return (CUDARenderer.cu:339)
	if (r != B2R_OK) std::cout << "!CUDARenderer::Render > " << b2r_last_error(m_pContext) << "!\n";
}

void CUDARenderer::Present()   // CUDARenderer.cu:383-395
{
	if (!m_pContext || !m_WindowHelper.pBackBufferPixels) return;
	b2r_read_framebuffer(m_pContext, m_WindowHelper.pBackBufferPixels);
}

void CUDARenderer::RenderAuto(const SceneManager& sm, const CUDATextureManager& tm, const Camera* pCamera)   // CUDARenderer.cu:357-371
{
	EnterValidRenderingState();
	Render(sm, tm, pCamera);
	Present();
}

unsigned int CUDARenderer::GetTotalNumVisibleTriangles() const
{
	uint32_t total = 0, visible = 0;
	if (m_pContext) b2r_get_stats(m_pContext, &total, &visible);
	return visible;
}
void CUDARenderer::DrawTextureImpl(const char* tp, int global, bool stretched, SampleState sampleState)   // CUDARenderer.cu:1968-2012
{
	if (!m_pContext) return;
	CUDATexture texture{};
	if (!texture.Create(tp)) { std::cout << "!CUDARenderer::DrawTexture > cannot load " << (tp ? tp : "(null)") << "!\n"; return; }
	EnterValidRenderingState();
	const float clear[3] = { m_ClearColour.r, m_ClearColour.g, m_ClearColour.b };
	if (b2r_draw_texture(m_pContext, texture.GetPixels().data(), texture.GetWidth(), texture.GetHeight(), global, stretched ? 1 : 0, (int)sampleState, clear) != B2R_OK)
	{
		std::cout << "!CUDARenderer::DrawTexture > " << b2r_last_error(m_pContext) << "!\n";
		return;
	}
	Present();
}
void CUDARenderer::DrawTexture(char* tp) { DrawTextureImpl(tp, 0, true, SampleState::Point); }
void CUDARenderer::DrawTextureGlobal(char* tp, bool isStretchedToWindow, SampleState sampleState) { DrawTextureImpl(tp, 1, isStretchedToWindow, sampleState); }
void CUDABenchMarker::StartTimer() { if (m_pRenderer) m_pRenderer->StartTimer(); }
float CUDABenchMarker::StopTimer() { return m_pRenderer ? m_pRenderer->StopTimer() : 0.f; }
void CUDARenderer::DisplayGPUSpecs(int) {}
void CUDARenderer::KernelWarmUp() {}
void CUDARenderer::StartTimer() { if (m_pContext) b2r_set_option(m_pContext, B2R_OPT_STAGE_TIMERS, 1); }
float CUDARenderer::StopTimer()
{
	float ms[B2R_NUM_STAGES] = {};
	if (m_pContext && b2r_get_stage_ms(m_pContext, ms, B2R_NUM_STAGES) == B2R_OK) return ms[B2R_STAGE_TOTAL];
	return 0.f;
}

// =========================================================================== Elite::Renderer
unsigned int Elite::Renderer::DefaultWidth = 640, Elite::Renderer::DefaultHeight = 480;
static WindowHelper ownedWindow(SDL_Window* pWindow, std::vector<unsigned int>& pixels)
{
	pixels.assign((size_t)Elite::Renderer::DefaultWidth * Elite::Renderer::DefaultHeight, 0u);
	WindowHelper h{};
	h.pWindow = pWindow; h.pBackBufferPixels = pixels.data();
	h.Resolution.Width = Elite::Renderer::DefaultWidth; h.Resolution.Height = Elite::Renderer::DefaultHeight;
	return h;
}
Elite::Renderer::Renderer(SDL_Window* pWindow) : m_OwnPixels{}, m_OwnHelper{ ownedWindow(pWindow, m_OwnPixels) }, m_WindowHelper{ m_OwnHelper }, m_Renderer{ m_OwnHelper } {}
Elite::Renderer::Renderer(const WindowHelper& windowHelper) : m_OwnPixels{}, m_OwnHelper{}, m_WindowHelper{ windowHelper }, m_Renderer{ windowHelper } {}
Elite::Renderer::~Renderer() = default;
void Elite::Renderer::LoadScene(const SceneGraph* pSceneGraph, const CUDATextureManager& tm) { m_pTextures = &tm; m_Renderer.LoadScene(pSceneGraph, tm); }
void Elite::Renderer::Render(const SceneManager& sm)   // ERenderer.cpp:42-164: same entry point, B200 path behind it
{
	if (!m_pCamera || !m_pTextures) return;
	m_Renderer.RenderAuto(sm, *m_pTextures, m_pCamera);
}
bool Elite::Renderer::SaveBackbufferToImage() const   // ERenderer.cpp:174-177 (SDL_SaveBMP replaced by a plain 32-bit BMP writer)
{
	const unsigned int w = m_WindowHelper.Resolution.Width, h = m_WindowHelper.Resolution.Height;
	if (!m_WindowHelper.pBackBufferPixels) return false;
	FILE* f = fopen("BackbufferRender.bmp", "wb");
	if (!f) return false;
	const uint32_t dataSize = w * h * 4, fileSize = 54 + dataSize;
	unsigned char hdr[54] = { 'B', 'M' };
	auto put32 = [&](int off, uint32_t v) { hdr[off] = v & 0xFF; hdr[off + 1] = (v >> 8) & 0xFF; hdr[off + 2] = (v >> 16) & 0xFF; hdr[off + 3] = (v >> 24) & 0xFF; };
	put32(2, fileSize); put32(10, 54); put32(14, 40); put32(18, w); put32(22, (uint32_t)(-(int32_t)h)); hdr[26] = 1; hdr[28] = 32; put32(34, dataSize);
	fwrite(hdr, 1, 54, f);
	fwrite(m_WindowHelper.pBackBufferPixels, 4, (size_t)w * h, f);   // 0x00RRGGBB words are BGRA bytes, BMP's native order
	fclose(f);
	return true;
}

// =========================================================================== C shims for the Python harness
// (tests drive the veneer's camera and OBJ parser through these; they are not part of include/b2r.h)
extern "C" {

// Camera(pos, fov); SetAspectRatio; SetNearFarValues; Pitch; Yaw  (the call order of main.cpp:421-425 / SURVEY config 5)
void b2r_host_camera(const float pos[3], float fovDeg, float width, float height, float nearZ, float farZ, float pitchDeg, float yawDeg,
	float* viewProj16, float* view16, float* proj16, float* camPos3, float* camFwd3)
{
	Camera cam{ FPoint3{ pos[0], pos[1], pos[2] }, fovDeg };
	cam.SetAspectRatio(width, height);
	cam.SetNearFarValues(nearZ, farZ);
	if (pitchDeg != 0.f) cam.Pitch(pitchDeg);
	if (yawDeg != 0.f) cam.Yaw(yawDeg);
	const FMatrix4 view = cam.GetViewMatrix();
	const FMatrix4 vp = cam.GetProjectionMatrix() * view;
	memcpy(viewProj16, vp.data, 64); memcpy(view16, view.data, 64); memcpy(proj16, cam.GetProjectionMatrix().data, 64);
	camPos3[0] = cam.GetPosition().x; camPos3[1] = cam.GetPosition().y; camPos3[2] = cam.GetPosition().z;
	camFwd3[0] = cam.GetForward().x; camFwd3[1] = cam.GetForward().y; camFwd3[2] = cam.GetForward().z;
}

// world matrix the renderer holds for a Mesh constructed with `pos` (pointer constructor) after LoadScene
void b2r_host_mesh_world(const float pos[3], int vectorCtor, float* world16)
{
	std::vector<IVertex> v; std::vector<unsigned int> i;
	Mesh* m = vectorCtor ? new Mesh{ v, (short)sizeof(IVertex), (short)0, i, PrimitiveTopology::TriangleList, FPoint3{ pos[0], pos[1], pos[2] } }
	                     : new Mesh{ nullptr, 0u, (short)sizeof(IVertex), (short)0, nullptr, 0u, PrimitiveTopology::TriangleList, FPoint3{ pos[0], pos[1], pos[2] } };
	FMatrix4 world = m->GetWorldConst();
	world.data[3][0] += m->GetPositionConst().x; world.data[3][1] += m->GetPositionConst().y; world.data[3][2] += m->GetPositionConst().z;
	memcpy(world16, world.data, 64);
	delete m;
}

struct b2r_host_obj { std::vector<IVertex> v; std::vector<unsigned int> idx; short vertexType = 0; };
b2r_host_obj* b2r_host_obj_parse(const char* path, int invertY, int pointerFlavour)
try
{
	if (!path) return nullptr;
	ObjParser parser;
	if (!parser.OpenFile(path)) return nullptr;
	parser.SetInvertYAxis(invertY != 0);
	std::unique_ptr<b2r_host_obj> o(new b2r_host_obj);
	if (pointerFlavour)
	{
		IVertex* pv = nullptr; unsigned int nv = 0, ni = 0; unsigned int* pi = nullptr;
		parser.ReadFromObjFile(pv, nv, pi, ni, o->vertexType);
		o->v.assign(pv, pv + nv); o->idx.assign(pi, pi + ni);
		delete[] pv; delete[] pi;
	}
	else
		parser.ReadFromObjFile(o->v, o->idx, o->vertexType);
	return o.release();
}
catch (...) { return nullptr; }      // malformed file (std::stof / stoi / substr throw): no exception crosses the C boundary
unsigned int b2r_host_obj_num_vertices(const b2r_host_obj* o) { return (unsigned int)o->v.size(); }
unsigned int b2r_host_obj_num_indices(const b2r_host_obj* o) { return (unsigned int)o->idx.size(); }
int b2r_host_obj_vertex_type(const b2r_host_obj* o) { return o->vertexType; }
const void* b2r_host_obj_vertices(const b2r_host_obj* o) { return o->v.data(); }
const unsigned int* b2r_host_obj_indices(const b2r_host_obj* o) { return o->idx.data(); }
void b2r_host_obj_free(b2r_host_obj* o) { delete o; }

int b2r_host_png_decode(const char* path, unsigned int* w, unsigned int* h, unsigned int* outPixels, unsigned long long capacity)
try
{
	std::vector<unsigned int> px; unsigned int ww = 0, hh = 0;
	if (!path || !w || !h || !DecodePNG(path, px, ww, hh)) return -1;
	*w = ww; *h = hh;
	if (outPixels && capacity >= px.size()) memcpy(outPixels, px.data(), px.size() * 4);
	return 0;
}
catch (...) { return -1; }

}
