// TEST INFRASTRUCTURE (oracle) -- not part of the product; only tests/, smoke() and
// bench.py's cpu_baseline leg may use anything under oracle/.
//
// Scalar restatement of the reference's header-only math (CustomCUDARasterizer/
// Vector3.h, Vector.h, Matrix3.h, Matrix4.h, Camera.cpp) in two flavours:
//
//  * host flavour  (h_*): what the reference's HOST code does when compiled for x86-64 -
//    plain IEEE single operations in source order, no fused multiply-add.
//  * device flavour (d_*): what the reference's DEVICE code does after nvcc 12.9 / ptxas
//    for sm_100a contracted it.  The contraction pattern was read from the SASS of
//    oracle/_ref/obj/CUDARenderer.o (see DESIGN.md "float semantics"):
//        a*b + c*d + e*f   ->  fma(e,f, fma(a,b, rn(c*d)))
//        a*b - c*d         ->  fma(a,b, -rn(c*d))
//        (a*b + c*d + e*f) + g  ->  rn(fma(e,f, fma(a,b, rn(c*d))) + g)
//    Division, sqrt and reciprocal are IEEE correctly rounded on both sides
//    (div.rn / sqrt.rn / rcp.rn in the reference build; no fast-math).
//
// Compile with -ffp-contract=off so that nothing here is fused implicitly; every fused
// operation is an explicit fmaf().
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

// data[col][row], exactly like Matrix4.h:17 / Matrix3.h:17
struct M4 { float d[4][4]; float operator()(int r, int c) const { return d[c][r]; } float& operator()(int r, int c) { return d[c][r]; } };
struct M3 { float d[3][3]; float operator()(int r, int c) const { return d[c][r]; } float& operator()(int r, int c) { return d[c][r]; } };

inline M4 Identity4()
{
	M4 m; std::memset(&m, 0, sizeof m);
	m.d[0][0] = m.d[1][1] = m.d[2][2] = m.d[3][3] = 1.f;
	return m;
}

// ---------------------------------------------------------------- host flavour
// Vector3.h:150-161
inline float h_Dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 h_Cross(const V3& a, const V3& b)
{
	return V3{ a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
}
inline V3 h_Scale(const V3& a, float s) { return V3{ a.x * s, a.y * s, a.z * s }; }
inline V3 h_Add(const V3& a, const V3& b) { return V3{ a.x + b.x, a.y + b.y, a.z + b.z }; }
inline V3 h_Sub(const V3& a, const V3& b) { return V3{ a.x - b.x, a.y - b.y, a.z - b.z }; }
// Vector.h:42-56
inline void h_Normalize(V3& v)
{
	const float m = std::sqrt(h_Dot(v, v));
	if (m == 0.f) { v = V3{ 0.f, 0.f, 0.f }; return; }
	const float invM = 1.f / m;
	v = h_Scale(v, invM);
}
// Matrix4.h:120-143
inline M4 h_Mul(const M4& lm, const M4& rm)
{
	M4 out;
	for (int r = 0; r < 4; ++r)
		for (int c = 0; c < 4; ++c)
			out(r, c) = lm(r, 0) * rm(0, c) + lm(r, 1) * rm(1, c) + lm(r, 2) * rm(2, c) + lm(r, 3) * rm(3, c);
	return out;
}
// Matrix3.h:126-133
inline V3 h_Mul(const M3& m, const V3& v)
{
	return V3{
		m(0, 0) * v.x + m(0, 1) * v.y + m(0, 2) * v.z,
		m(1, 0) * v.x + m(1, 1) * v.y + m(1, 2) * v.z,
		m(2, 0) * v.x + m(2, 1) * v.y + m(2, 2) * v.z };
}
// Matrix4.h:328-363
inline M4 h_Inverse(const M4& m)
{
	const V3 a{ m.d[0][0], m.d[0][1], m.d[0][2] };
	const V3 b{ m.d[1][0], m.d[1][1], m.d[1][2] };
	const V3 c{ m.d[2][0], m.d[2][1], m.d[2][2] };
	const V3 d{ m.d[3][0], m.d[3][1], m.d[3][2] };
	const float x = m(3, 0), y = m(3, 1), z = m(3, 2), w = m(3, 3);
	V3 s = h_Cross(a, b);
	V3 t = h_Cross(c, d);
	V3 u = h_Sub(h_Scale(a, y), h_Scale(b, x));
	V3 v = h_Sub(h_Scale(c, w), h_Scale(d, z));
	const float det = h_Dot(s, v) + h_Dot(t, u);
	const float invDet = 1.0f / det;
	s = h_Scale(s, invDet); t = h_Scale(t, invDet); u = h_Scale(u, invDet); v = h_Scale(v, invDet);
	const V3 r0 = h_Add(h_Cross(b, v), h_Scale(t, y));
	const V3 r1 = h_Sub(h_Cross(v, a), h_Scale(t, x));
	const V3 r2 = h_Add(h_Cross(d, u), h_Scale(s, w));
	const V3 r3 = h_Sub(h_Cross(u, c), h_Scale(s, z));
	M4 o;
	o(0, 0) = r0.x; o(0, 1) = r0.y; o(0, 2) = r0.z; o(0, 3) = -h_Dot(b, t);
	o(1, 0) = r1.x; o(1, 1) = r1.y; o(1, 2) = r1.z; o(1, 3) = h_Dot(a, t);
	o(2, 0) = r2.x; o(2, 1) = r2.y; o(2, 2) = r2.z; o(2, 3) = -h_Dot(d, s);
	o(3, 0) = r3.x; o(3, 1) = r3.y; o(3, 2) = r3.z; o(3, 3) = h_Dot(c, s);
	return o;
}

// ---------------------------------------------------------------- device flavour
inline float d_sum3(float a, float b, float c, float d, float e, float f) // a*b + c*d + e*f
{
	return fmaf(e, f, fmaf(a, b, c * d));
}
inline float d_diff2(float a, float b, float c, float d) // a*b - c*d
{
	return fmaf(a, b, -(c * d));
}
inline float d_Dot(const V3& a, const V3& b) { return d_sum3(a.x, b.x, a.y, b.y, a.z, b.z); }
inline V3 d_Cross(const V3& a, const V3& b)
{
	return V3{ d_diff2(a.y, b.z, a.z, b.y), d_diff2(a.z, b.x, a.x, b.z), d_diff2(a.x, b.y, a.y, b.x) };
}
inline float d_Cross2(const V2& a, const V2& b) { return d_diff2(a.x, b.y, a.y, b.x); }
inline void d_Normalize(V3& v)
{
	const float m = std::sqrt(d_Dot(v, v));
	if (m == 0.f) { v = V3{ 0.f, 0.f, 0.f }; return; }
	const float invM = 1.f / m;
	v = V3{ v.x * invM, v.y * invM, v.z * invM };
}
// Matrix4.h:158-166 (point: w assumed 1)
inline V4 d_MulPoint(const M4& m, const V3& p)
{
	return V4{
		d_sum3(m(0, 0), p.x, m(0, 1), p.y, m(0, 2), p.z) + m(0, 3),
		d_sum3(m(1, 0), p.x, m(1, 1), p.y, m(1, 2), p.z) + m(1, 3),
		d_sum3(m(2, 0), p.x, m(2, 1), p.y, m(2, 2), p.z) + m(2, 3),
		d_sum3(m(3, 0), p.x, m(3, 1), p.y, m(3, 2), p.z) + m(3, 3) };
}
inline V3 d_Mul(const M3& m, const V3& v)
{
	return V3{
		d_sum3(m(0, 0), v.x, m(0, 1), v.y, m(0, 2), v.z),
		d_sum3(m(1, 0), v.x, m(1, 1), v.y, m(1, 2), v.z),
		d_sum3(m(2, 0), v.x, m(2, 1), v.y, m(2, 2), v.z) };
}

// cvt.rzi.s32.f32 : round toward zero, saturating, NaN -> 0
inline int32_t cvt_rzi_s32(float f)
{
	if (f != f) return 0;
	if (f >= 2147483648.f) return INT32_MAX;
	if (f <= -2147483648.f) return INT32_MIN;
	return (int32_t)f;
}
// cvt.rzi.u32.f32 : round toward zero, saturating, negatives and NaN -> 0
inline uint32_t cvt_rzi_u32(float f)
{
	if (!(f > 0.f)) return 0u;
	if (f >= 4294967296.f) return UINT32_MAX;
	return (uint32_t)f;
}


// ---- Camera.cpp:6-37, 45-50, 67-81, 97-117
struct Camera
{
	float fov, aspect = 0, nearZ = 0.1f, farZ = 100.f;
	M4 proj, lookAt;
	explicit Camera(const float pos[3], float fovDeg)
	{
		const float radians = fovDeg * (float(3.14159265358979323846) / 180.f);   // MathUtilities.h:75-78
		fov = tanf(radians / 2.f);
		proj = Identity4();
		lookAt = Identity4();
		lookAt.d[3][0] = pos[0]; lookAt.d[3][1] = pos[1]; lookAt.d[3][2] = pos[2];
	}
	V3& Forward() { return *reinterpret_cast<V3*>(&lookAt.d[2][0]); }
	V3& Right() { return *reinterpret_cast<V3*>(&lookAt.d[0][0]); }
	V3& Up() { return *reinterpret_cast<V3*>(&lookAt.d[1][0]); }
	void RecalculateProjection()
	{
		const float A = farZ / (nearZ - farZ);
		const float B = (farZ * nearZ) / (nearZ - farZ);
		const float value = 1.f / (aspect * fov);
		const float invFov = 1.f / fov;
		proj.d[0][0] = value;
		proj.d[1][1] = invFov;
		proj.d[2][2] = A;
		proj.d[2][3] = -1.f;
		proj.d[3][2] = B;
	}
	void SetAspectRatio(float w, float h) { aspect = w / h; RecalculateProjection(); }
	void SetNearFar(float n, float f) { nearZ = n; farZ = f; RecalculateProjection(); }
	void RecalculateVectors()
	{
		const V3 worldUp{ 0.f, 1.f, 0.f };
		h_Normalize(Forward());
		Right() = h_Cross(worldUp, Forward());
		h_Normalize(Right());
		Up() = h_Cross(Forward(), Right());
		h_Normalize(Up());
	}
	void Pitch(float degrees)   // Matrix3.h:320-329
	{
		const float radians = degrees * (float(3.14159265358979323846) / 180.f);
		const float c = (float)cos(radians), s = (float)sin(radians);
		M3 m; m(0, 0) = 1; m(0, 1) = 0; m(0, 2) = 0; m(1, 0) = 0; m(1, 1) = c; m(1, 2) = -s; m(2, 0) = 0; m(2, 1) = s; m(2, 2) = c;
		Forward() = h_Mul(m, Forward());
		RecalculateVectors();
	}
	void Yaw(float degrees)     // Matrix3.h:332-341
	{
		const float radians = degrees * (float(3.14159265358979323846) / 180.f);
		const float c = (float)cos(radians), s = (float)sin(radians);
		M3 m; m(0, 0) = c; m(0, 1) = 0; m(0, 2) = s; m(1, 0) = 0; m(1, 1) = 1; m(1, 2) = 0; m(2, 0) = -s; m(2, 1) = 0; m(2, 2) = c;
		Forward() = h_Mul(m, Forward());
		RecalculateVectors();
	}
};

}
