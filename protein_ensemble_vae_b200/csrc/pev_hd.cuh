// Host/device math shared by the CUDA kernels and by tests/hostcheck (which compiles the very
// same per-thread bodies with g++ so their arithmetic can be checked against the oracle on a
// machine without a GPU).  No CUDA runtime calls in here.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PEV_HD __host__ __device__ __forceinline__
#else
#define PEV_HD inline
#endif

namespace pev {

struct v3 {
  float x, y, z;
};

PEV_HD v3 mk3(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
PEV_HD v3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }
PEV_HD void st3(float* p, v3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
PEV_HD v3 operator+(v3 a, v3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
PEV_HD v3 operator-(v3 a, v3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
PEV_HD v3 operator*(v3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
PEV_HD v3 operator*(float s, v3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
PEV_HD v3& operator+=(v3& a, v3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
PEV_HD v3& operator-=(v3& a, v3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
PEV_HD float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PEV_HD v3 cross(v3 a, v3 b) {
  return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
PEV_HD float norm(v3 a) { return sqrtf(dot(a, a)); }
PEV_HD v3 zero3() { return mk3(0.f, 0.f, 0.f); }

PEV_HD float sigmoidf_(float z) { return 1.0f / (1.0f + expf(-z)); }
PEV_HD float siluf_(float z) { return z * sigmoidf_(z); }
// d silu / dz = sigma (1 + z (1 - sigma))
PEV_HD float silu_gradf_(float z) {
  float s = sigmoidf_(z);
  return s * (1.0f + z * (1.0f - s));
}

// huber_loss, models/losses.py:311-316
PEV_HD float huberf_(float e, float delta) {
  float a = fabsf(e);
  return a < delta ? 0.5f * e * e : delta * (a - 0.5f * delta);
}
PEV_HD float huber_gradf_(float e, float delta) {
  float a = fabsf(e);
  return a < delta ? e : (e > 0.f ? delta : (e < 0.f ? -delta : 0.f));
}

PEV_HD float signf_(float a) { return a > 0.f ? 1.f : (a < 0.f ? -1.f : 0.f); }

// y = a / (|a| + eps); returns g_a given g_y.  torch.norm's sub-gradient at 0 is 0.
PEV_HD v3 normalize_eps_bwd(v3 a, float len, float eps, v3 gy) {
  float inv = 1.0f / (len + eps);
  v3 g = gy * inv;
  if (len > 0.f) {
    float k = dot(gy, a) * inv * inv / len;
    g -= a * k;
  }
  return g;
}

// ------------------------------------------------------------------------------------------
// torsion p0-p1-p2-p3 as (sin, cos); _dihedral_from_four, models/losses.py:158-232.
struct Dihedral {
  float s, c;
  // saved for backward
  v3 b1, b2, b3, n1, n2, u1, u2;
  float l1, l2, craw, sgn;
  bool ok;
};

PEV_HD Dihedral dihedral_fwd(v3 p0, v3 p1, v3 p2, v3 p3) {
  const float eps = 1e-8f;
  Dihedral d;
  d.b1 = p1 - p0;
  d.b2 = p2 - p1;
  d.b3 = p3 - p2;
  d.n1 = cross(d.b1, d.b2);
  d.n2 = cross(d.b2, d.b3);
  d.l1 = norm(d.n1);
  d.l2 = norm(d.n2);
  d.ok = (d.l1 > eps) && (d.l2 > eps);
  d.s = 0.f;
  d.c = 1.f;
  d.craw = 1.f;
  d.sgn = 0.f;
  d.u1 = zero3();
  d.u2 = zero3();
  if (d.ok) {
    d.u1 = d.n1 * (1.0f / (d.l1 + eps));
    d.u2 = d.n2 * (1.0f / (d.l2 + eps));
    v3 ub = d.b2 * (1.0f / (norm(d.b2) + eps));
    d.craw = dot(d.u1, d.u2);
    float c = fminf(fmaxf(d.craw, -1.0f + eps), 1.0f - eps);   // no-op bounds in fp32 (F10 quirk list)
    d.sgn = signf_(dot(cross(d.u1, d.u2), ub));
    d.c = c;
    d.s = d.sgn * sqrtf(1.0f - c * c + eps);
  }
  return d;
}

// gradient of (gs * sin + gc * cos) w.r.t. the four points, following autograd through the
// reference's formula: both outputs depend on the points only through cos = clamp(u1.u2);
// sign() and the validity test carry no gradient.
PEV_HD void dihedral_bwd(const Dihedral& d, float gs, float gc, v3& g0, v3& g1, v3& g2, v3& g3) {
  const float eps = 1e-8f;
  g0 = g1 = g2 = g3 = zero3();
  if (!d.ok) return;
  float gcos = gc;
  gcos += gs * d.sgn * (-d.c) / sqrtf(1.0f - d.c * d.c + eps);
  if (!(d.craw >= -1.0f + eps && d.craw <= 1.0f - eps)) gcos = 0.f;     // clamp passes grad inside only
  v3 gn1 = normalize_eps_bwd(d.n1, d.l1, eps, d.u2 * gcos);
  v3 gn2 = normalize_eps_bwd(d.n2, d.l2, eps, d.u1 * gcos);
  // n1 = b1 x b2, n2 = b2 x b3
  v3 gb1 = cross(d.b2, gn1);
  v3 gb2 = cross(gn1, d.b1) + cross(d.b3, gn2);
  v3 gb3 = cross(gn2, d.b2);
  g0 = zero3() - gb1;
  g1 = gb1 - gb2;
  g2 = gb2 - gb3;
  g3 = gb3;
}

// ------------------------------------------------------------------------------------------
// angle at vertex b (radians); _angle_cos + acos, models/losses.py:358-368, :383.
struct Angle {
  float theta, cosv, lu, lv;
  v3 u, v;
};
PEV_HD Angle angle_fwd(v3 a, v3 b, v3 c) {
  const float eps = 1e-8f;
  Angle r;
  r.u = a - b;
  r.v = c - b;
  r.lu = norm(r.u);
  r.lv = norm(r.v);
  v3 un = r.u * (1.0f / (r.lu + eps));
  v3 vn = r.v * (1.0f / (r.lv + eps));
  r.cosv = fminf(fmaxf(dot(un, vn), -1.0f), 1.0f);
  r.theta = acosf(r.cosv);
  return r;
}
PEV_HD void angle_bwd(const Angle& r, float gtheta, v3& ga, v3& gb, v3& gc) {
  const float eps = 1e-8f;
  // d acos / d cos = -1 / sqrt(1 - cos^2) is infinite where the clamp of _angle_cos is active (collinear atoms: the reference
  // back-propagates inf / NaN there, models/losses.py:366,383; at random init one of ~2e5 angles per batch gets there within
  // a few dozen steps and poisons every parameter).  Here the clamped cosine passes no gradient.
  const float s2 = 1.0f - r.cosv * r.cosv;
  float gcos = s2 > 1e-12f ? -gtheta / sqrtf(s2) : 0.0f;
  v3 un = r.u * (1.0f / (r.lu + eps));
  v3 vn = r.v * (1.0f / (r.lv + eps));
  v3 gu = normalize_eps_bwd(r.u, r.lu, eps, vn * gcos);
  v3 gv = normalize_eps_bwd(r.v, r.lv, eps, un * gcos);
  ga = gu;
  gc = gv;
  gb = zero3() - gu - gv;
}

// ------------------------------------------------------------------------------------------
// Ramachandran penalty from (sin,cos) slots; ramachandran_loss, models/losses.py:88-129.
PEV_HD float rama_penalty(float sphi, float cphi, float spsi, float cpsi, float* dphi, float* dpsi) {
  float phi = atan2f(sphi, cphi), psi = atan2f(spsi, cpsi);
  const float P0[4] = {-1.05f, -2.09f, 1.05f, -1.31f};
  const float S0[4] = {-0.79f, 2.09f, 0.79f, 2.53f};
  const float WD[4] = {0.6f, 0.9f, 0.6f, 0.5f};
  float best = -1.f;
  int arg = 0;
  for (int k = 0; k < 4; ++k) {
    float a = phi - P0[k], b = psi - S0[k];
    float g = expf(-(a * a / WD[k] + b * b / WD[k]));
    if (g > best) { best = g; arg = k; }
  }
  if (dphi) {
    *dphi = best * (2.0f * (phi - P0[arg]) / WD[arg]);      // d(1 - best)/dphi
    *dpsi = best * (2.0f * (psi - S0[arg]) / WD[arg]);
  }
  float pen = 1.0f - best;
  if (phi > 0.f && psi < 0.f) pen += 5.0f;
  return pen;
}

// omega_trans_loss penalty, models/losses.py:144-150.  *dom = d pen / d omega.
PEV_HD float omega_penalty(float som, float com, float* dom) {
  const float PI_F = 3.14159265358979323846f;
  float om = atan2f(som, com);
  float pen = 2.0f * (1.0f - cosf(om - PI_F));
  float wrapped = atan2f(sinf(om), cosf(om));
  if (fabsf(wrapped) < 0.5f) pen += 3.0f;
  if (dom) *dom = 2.0f * sinf(om - PI_F);
  return pen;
}

// d atan2(s, c): (ds, dc) coefficients; torch gives 0/0 = NaN at the origin but those slots are
// constants in the reference (zero-filled, never connected to coordinates), so 0 is returned.
PEV_HD void atan2_grad(float s, float c, float g, float* gs, float* gc) {
  float r2 = s * s + c * c;
  if (r2 > 0.f) {
    *gs = g * c / r2;
    *gc = -g * s / r2;
  } else {
    *gs = 0.f;
    *gc = 0.f;
  }
}

PEV_HD bool finitef_(float a) { return fabsf(a) <= 3.402823466e38f; }   // false for NaN and +-inf

}  // namespace pev
