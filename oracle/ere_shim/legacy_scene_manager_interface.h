// TEST INFRASTRUCTURE (oracle) -- the SceneManager interface ERenderer.cpp was written against: the current
// SceneManager.h (SceneManager.h:9-36) plus IsDepthColour(), which the current class expresses as
// GetVisualisationState() == VisualisationState::Depth.  OUR header-only class: no EventManager, no SDL input.
#pragma once
#include <vector>
#include "CullingMode.h"
#include "SampleState.h"
#include "VisualisationState.h"

class SceneGraph;
class SceneManager final
{
public:
	SceneManager() = default;

	void AddSceneGraph(SceneGraph* pSceneGraph) { m_pSceneGraphs.push_back(pSceneGraph); }
	SceneGraph* GetSceneGraph() const { return m_pSceneGraphs.empty() ? nullptr : m_pSceneGraphs[m_Index]; }

	VisualisationState GetVisualisationState() const { return m_Visualisation; }
	SampleState GetSampleState() const { return m_SampleState; }
	CullingMode GetCullingMode() const { return m_CullingMode; }
	bool IsDepthColour() const { return m_Visualisation == VisualisationState::Depth; }

	void SetVisualisationState(VisualisationState v) { m_Visualisation = v; }
	void SetSampleState(SampleState s) { m_SampleState = s; }
	void SetCullingMode(CullingMode c) { m_CullingMode = c; }

private:
	int m_Index{};
	VisualisationState m_Visualisation{};
	std::vector<SceneGraph*> m_pSceneGraphs;
	SampleState m_SampleState{};
	CullingMode m_CullingMode{};
};
