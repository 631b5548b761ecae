// TEST INFRASTRUCTURE (oracle) -- the Mesh interface ERenderer.cpp was written against (it predates the reference's
// current Mesh.h): vector-backed buffers, GetVertexAmount / GetIndexAmount, per-mesh CPU `Textures`, GetWorldMatrix.
// OUR header, reconstructed from the call sites in CustomCUDARasterizer/ERenderer.cpp:59-69 and SceneGraph.cpp:6-39;
// the world-matrix convention is the current Mesh's (Mesh.cpp:15-31: identity with the position added to row 3).
#pragma once
#include <vector>
#include "Vertex.h"
#include "PrimitiveTopology.h"
#include "Textures.h"

class Mesh
{
public:
	Mesh(std::vector<IVertex> vertexBuffer, std::vector<unsigned int> indexBuffer, PrimitiveTopology topology, const FPoint3& pos, const Textures& textures)
		: m_VertexBuffer{ std::move(vertexBuffer) }, m_IndexBuffer{ std::move(indexBuffer) }, m_Topology{ topology }
		, m_Textures{ textures }, m_WorldSpace{ FMatrix4::Identity() }
	{
		reinterpret_cast<FPoint3&>(m_WorldSpace[3][0]) += reinterpret_cast<const FVector3&>(pos);
	}
	virtual ~Mesh() = default;

	const std::vector<IVertex>& GetVertexBuffer() const { return m_VertexBuffer; }
	unsigned int GetVertexAmount() const { return (unsigned int)m_VertexBuffer.size(); }
	const std::vector<unsigned int>& GetIndexBuffer() const { return m_IndexBuffer; }
	unsigned int GetIndexAmount() const { return (unsigned int)m_IndexBuffer.size(); }
	const Textures& GetTextures() const { return m_Textures; }
	PrimitiveTopology GetTopology() const { return m_Topology; }
	const FMatrix4& GetWorldMatrix() const { return m_WorldSpace; }
	FMatrix4& GetWorld() { return m_WorldSpace; }

private:
	std::vector<IVertex> m_VertexBuffer;
	std::vector<unsigned int> m_IndexBuffer;
	PrimitiveTopology m_Topology;
	Textures m_Textures;
	FMatrix4 m_WorldSpace;
};
