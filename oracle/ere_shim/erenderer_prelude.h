// TEST INFRASTRUCTURE (oracle) -- OUR replacement of the reference's precompiled header for ONE purpose: compiling the
// reference's CPU renderer, CustomCUDARasterizer/ERenderer.cpp, unmodified (oracle/build_ref_erenderer.sh).
// The reference's own PCH.h pulls in the CUDA renderer and the current Mesh/SceneManager interfaces, which ERenderer.cpp
// predates (it still calls Mesh::GetVertexAmount / GetTextures / GetWorldMatrix and SceneManager::IsDepthColour - the
// file is excluded from the reference's project, vcxproj:126-155).  This header provides only what the translation
// unit needs: the standard headers the MSVC build got transitively, and the reference's own math / colour headers.
#pragma once
#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <iostream>
#include <math.h>
#include <stdlib.h>
#include <type_traits>
#include <vector>

// the reference's headers, as they lie
#include "GPUHelpers.h"
#include "Math.h"
#include "MathUtilities.h"
#include "RGBColor.h"
#include "RGBRaw.h"

// ERenderer.cpp:509 calls unqualified abs() on a float; with MSVC's headers that is the float overload
static_assert(std::is_same<decltype(abs(1.0f)), float>::value, "abs(float) must not truncate to int");

constexpr int NumTextures = 4;
