// Minimal host-side math types of the scene-model boundary.
//
// The reference exposes its own header-only math library (Math.h, Vector*.h, Point*.h,
// Matrix*.h; ~2.5 kLoC) through the signatures of Camera, Mesh, CUDARenderer ... .  Only
// the handful of PODs those signatures use are provided here, with the same names, the
// same memory layout (matrices are data[col][row], Matrix4.h:17) and host arithmetic in
// the same order (plain IEEE single precision, no fused multiply-add: build with
// -ffp-contract=off), because the matrices computed here feed discrete decisions on the
// device (bounding boxes, bins).
#pragma once
#include <cmath>
#include <cstdint>

struct FVector2 { float x = 0.f, y = 0.f; FVector2() = default; FVector2(float x_, float y_) : x(x_), y(y_) {} };

struct FVector3
{
	float x = 0.f, y = 0.f, z = 0.f;
	FVector3() = default;
	FVector3(float x_, float y_, float z_ = 0.f) : x(x_), y(y_), z(z_) {}
	FVector3 operator+(const FVector3& o) const { return { x + o.x, y + o.y, z + o.z }; }
	FVector3 operator-(const FVector3& o) const { return { x - o.x, y - o.y, z - o.z }; }
	FVector3 operator*(float s) const { return { x * s, y * s, z * s }; }
	FVector3& operator+=(const FVector3& o) { x += o.x; y += o.y; z += o.z; return *this; }
};

struct FPoint3
{
	float x = 0.f, y = 0.f, z = 0.f;
	FPoint3() = default;
	FPoint3(float x_, float y_, float z_ = 1.f) : x(x_), y(y_), z(z_) {}
	FVector3 operator-(const FPoint3& o) const { return { x - o.x, y - o.y, z - o.z }; }
	FPoint3& operator+=(const FVector3& v) { x += v.x; y += v.y; z += v.z; return *this; }
};

struct FPoint4 { float x = 0.f, y = 0.f, z = 0.f, w = 0.f; };
struct IPoint2 { int x = 0, y = 0; IPoint2() = default; IPoint2(int x_, int y_) : x(x_), y(y_) {} };

struct RGBColor { float r = 0.f, g = 0.f, b = 0.f; RGBColor() = default; RGBColor(float r_, float g_, float b_) : r(r_), g(g_), b(b_) {} };

// Vector3.h:150-161 (host evaluation order)
inline float Dot(const FVector3& a, const FVector3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline FVector3 Cross(const FVector3& a, const FVector3& b)
{
	return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
}
inline float Cross(const FVector2& a, const FVector2& b) { return a.x * b.y - a.y * b.x; }
// Vector.h:42-63
inline float Normalize(FVector3& v)
{
	const float m = std::sqrt(Dot(v, v));
	if (m == 0.f) { v = FVector3{}; return m; }
	const float inv = 1.f / m;
	v = v * inv;
	return m;
}
inline FVector3 GetNormalized(const FVector3& v) { FVector3 c = v; Normalize(c); return c; }
// Vector.h:66-73
inline FVector3 Project(const FVector3& v, const FVector3& t) { return t * (Dot(v, t) / Dot(t, t)); }
inline FVector3 Reject(const FVector3& v, const FVector3& t) { return v - Project(v, t); }

struct FMatrix3
{
	float data[3][3] = {};     // data[col][row]
	float operator()(int r, int c) const { return data[c][r]; }
	float& operator()(int r, int c) { return data[c][r]; }
	// Matrix3.h:126-133
	FVector3 operator*(const FVector3& v) const
	{
		const FMatrix3& m = *this;
		return { m(0, 0) * v.x + m(0, 1) * v.y + m(0, 2) * v.z,
		         m(1, 0) * v.x + m(1, 1) * v.y + m(1, 2) * v.z,
		         m(2, 0) * v.x + m(2, 1) * v.y + m(2, 2) * v.z };
	}
};

struct FMatrix4
{
	float data[4][4] = {};     // data[col][row]
	float operator()(int r, int c) const { return data[c][r]; }
	float& operator()(int r, int c) { return data[c][r]; }
	static FMatrix4 Identity()
	{
		FMatrix4 m;
		m.data[0][0] = m.data[1][1] = m.data[2][2] = m.data[3][3] = 1.f;
		return m;
	}
	// Matrix4.h:120-143
	FMatrix4 operator*(const FMatrix4& rm) const
	{
		const FMatrix4& lm = *this;
		FMatrix4 o;
		for (int r = 0; r < 4; ++r)
			for (int c = 0; c < 4; ++c)
				o(r, c) = lm(r, 0) * rm(0, c) + lm(r, 1) * rm(1, c) + lm(r, 2) * rm(2, c) + lm(r, 3) * rm(3, c);
		return o;
	}
	FMatrix4& operator*=(const FMatrix4& m) { *this = *this * m; return *this; }
	explicit operator FMatrix3() const   // Matrix3.h:51-56
	{
		FMatrix3 m;
		for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) m.data[c][r] = data[c][r];
		return m;
	}
};

// Matrix3.h:320-341 (angle in radians)
inline FMatrix3 MakeRotationX(float t)
{
	const float c = (float)std::cos(t), s = (float)std::sin(t);
	FMatrix3 m;
	m(0, 0) = 1.f; m(1, 1) = c; m(1, 2) = -s; m(2, 1) = s; m(2, 2) = c;
	return m;
}
inline FMatrix3 MakeRotationY(float t)
{
	const float c = (float)std::cos(t), s = (float)std::sin(t);
	FMatrix3 m;
	m(0, 0) = c; m(0, 2) = s; m(1, 1) = 1.f; m(2, 0) = -s; m(2, 2) = c;
	return m;
}
inline FMatrix4 MakeRotationY4(float t)
{
	const FMatrix3 r = MakeRotationY(t);
	FMatrix4 m = FMatrix4::Identity();
	for (int c = 0; c < 3; ++c) for (int rr = 0; rr < 3; ++rr) m.data[c][rr] = r.data[c][rr];
	return m;
}

// Matrix4.h:328-363
inline FMatrix4 Inverse(const FMatrix4& m)
{
	const FVector3 a{ m.data[0][0], m.data[0][1], m.data[0][2] };
	const FVector3 b{ m.data[1][0], m.data[1][1], m.data[1][2] };
	const FVector3 c{ m.data[2][0], m.data[2][1], m.data[2][2] };
	const FVector3 d{ m.data[3][0], m.data[3][1], m.data[3][2] };
	const float x = m(3, 0), y = m(3, 1), z = m(3, 2), w = m(3, 3);
	FVector3 s = Cross(a, b), t = Cross(c, d);
	FVector3 u = a * y - b * x, v = c * w - d * z;
	const float det = Dot(s, v) + Dot(t, u);
	const float invDet = 1.0f / det;
	s = s * invDet; t = t * invDet; u = u * invDet; v = v * invDet;
	const FVector3 r0 = Cross(b, v) + t * y;
	const FVector3 r1 = Cross(v, a) - t * x;
	const FVector3 r2 = Cross(d, u) + s * w;
	const FVector3 r3 = Cross(u, c) - s * z;
	FMatrix4 o;
	o(0, 0) = r0.x; o(0, 1) = r0.y; o(0, 2) = r0.z; o(0, 3) = -Dot(b, t);
	o(1, 0) = r1.x; o(1, 1) = r1.y; o(1, 2) = r1.z; o(1, 3) = Dot(a, t);
	o(2, 0) = r2.x; o(2, 1) = r2.y; o(2, 2) = r2.z; o(2, 3) = -Dot(d, s);
	o(3, 0) = r3.x; o(3, 1) = r3.y; o(3, 2) = r3.z; o(3, 3) = Dot(c, s);
	return o;
}

// MathUtilities.h:75-78
inline float ToRadians(float degrees) { return degrees * (float(3.14159265358979323846) / 180.f); }

// MathUtilities.h:31-38 (ulp = 2)
inline bool AreEqual(float a, float b)
{
	return std::fabs(a - b) <= 1.1920929e-07f * std::fabs(a + b) * 2
		|| std::fabs(a - b) < 1.17549435e-38f;
}
