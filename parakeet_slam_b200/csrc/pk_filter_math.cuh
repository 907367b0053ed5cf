// fp64 per-landmark math shared by the fused measurement kernel (pk_measure.cu) and the
// single-call probe entry points (pk_probe.cu): probability_of_match and one EKF update.
#pragma once

#include <math.h>

#include "pk_common.cuh"

namespace pk {

// fp64 literals are materialised by two UMOV/IMAD.MOV each; coefficients placed in __constant__ memory are
// consumed directly as c[bank][offset] operands of DFMA / DADD.
__constant__ double kExpC[12] = {1.0 / 6.0,        0.5,
                                 1.0 / 120.0,      1.0 / 24.0,
                                 1.0 / 5040.0,     1.0 / 720.0,
                                 1.0 / 362880.0,   1.0 / 40320.0,
                                 1.0 / 39916800.0, 1.0 / 3628800.0,
                                 1.0 / 6227020800.0, 1.0 / 479001600.0};
__constant__ double kExpK[4] = {1.4426950408889634074, 6755399441055744.0, -6.93147180369123816490e-01,
                                -1.90821492927058770002e-10};
__constant__ double kAtanC[11] = {3.33333333333329318027e-01,  -1.99999999998764832476e-01, 1.42857142725034663711e-01,
                                  -1.11111104054623557880e-01, 9.09088713343650656196e-02,  -7.69187620504482999495e-02,
                                  6.66107313738753120669e-02,  -5.83357013379057348645e-02, 4.97687799461593236017e-02,
                                  -3.65315727442169155270e-02, 1.62858201153657823623e-02};
__constant__ double kAtanK[6] = {7.85398163397448278999e-01, 3.06161699786838301793e-17, 1.57079632679489655800e+00,
                                 6.12323399573676603587e-17, 3.14159265358979311600e+00, 1.22464679914735317720e-16};

// ---------------------------------------------------------------------------------------------
// Branch-free exp (fp64, <= 2 ulp, exact underflow behaviour).  Most arguments here are hundreds
// below zero (finding F3: the match / no-match decision IS the fp64 underflow of the likelihood), which
// is libm's slow path, and lanes diverge between fast and slow path.  n = rint(x/ln2), r = x - n ln2
// (two-constant Cody-Waite), degree-13 Taylor polynomial of e^r by Estrin's scheme, and the scaling by
// 2^n split in two exact powers so that the last multiply performs the single, correctly rounded
// step into the subnormal range (or to 0 / inf).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double pk_exp(double x) {
    double xc = (x < -1100.0) ? -1100.0 : x;  // (plain selects: fmin/fmax cost ~10 instructions each in fp64)
    xc = (xc > 1100.0) ? 1100.0 : xc;
    const double shift = kExpK[1];  // 1.5 * 2^52: adding it rounds to nearest integer
    const double t = fma(xc, kExpK[0], shift);
    const int n = __double2loint(t);
    const double fn = t - shift;
    double r = fma(fn, kExpK[2], xc);
    r = fma(fn, kExpK[3], r);
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double a0 = 1.0 + r;
    const double a1 = fma(r, kExpC[0], kExpC[1]);
    const double a2 = fma(r, kExpC[2], kExpC[3]);
    const double a3 = fma(r, kExpC[4], kExpC[5]);
    const double a4 = fma(r, kExpC[6], kExpC[7]);
    const double a5 = fma(r, kExpC[8], kExpC[9]);
    const double a6 = fma(r, kExpC[10], kExpC[11]);
    const double lo = fma(r2, a1, a0);                       // 1 + r + r^2 (1/2 + r/6)
    const double mid = fma(r2, a3, a2);                      // r^4 ( ... )
    const double hi = fma(r4, a6, fma(r2, a5, a4));          // r^8 ( ... )
    const double p = fma(r8, hi, fma(r4, mid, lo));
    const int n1 = n >> 1, n2 = n - n1;
    const double s1 = __hiloint2double((n1 + 1023) << 20, 0);
    const double s2 = __hiloint2double((n2 + 1023) << 20, 0);
    const double y = (p * s1) * s2;
    return (x != x) ? x : y;
}

// ---------------------------------------------------------------------------------------------
// Branch-free atan2 (fp64, <= 1.5 ulp).  libm's atan2 costs ~250 warp-instructions here because
// its quadrant / magnitude cases diverge across the lanes of a warp.  One division: with
// mn = min(|x|,|y|), mx = max(|x|,|y|), either t = mn/mx (t <= tan(pi/8)) or
// t = (mn - mx)/(mn + mx) and pi/4 is added, so |t| <= tan(pi/8) < 7/16 always, where the odd
// minimax polynomial of fdlibm's s_atan.c (aT[0..10], error < 1 ulp) applies.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double pk_atan2(double y, double x) {
    const double ax = fabs(x), ay = fabs(y);
    const bool steep = ay > ax;
    const double mx = steep ? ay : ax, mn = steep ? ax : ay;
    const bool big = mn > 0.41421356237309503 * mx;
    const double num = big ? mn - mx : mn;
    const double den = big ? mn + mx : mx;
    double t = num / den;
    if (den == 0.0) t = 0.0;  // atan2(+-0, +-0): math.atan2 gives +-0 / +-pi like the selects below
    const double z = t * t, w = z * z;
    const double s1 = z * fma(w, fma(w, fma(w, fma(w, fma(w, kAtanC[10], kAtanC[8]), kAtanC[6]), kAtanC[4]), kAtanC[2]),
                              kAtanC[0]);
    const double s2 = w * fma(w, fma(w, fma(w, fma(w, kAtanC[9], kAtanC[7]), kAtanC[5]), kAtanC[3]), kAtanC[1]);
    double r = t - t * (s1 + s2);
    if (big) r = kAtanK[0] + (r + kAtanK[1]);
    if (steep) r = kAtanK[2] - (r - kAtanK[3]);
    if (signbit(x)) r = kAtanK[4] - (r - kAtanK[5]);
    return copysign(r, y);
}

constexpr double kLog2Pi = 1.8378770664093453;  // math.log(2*pi)

// ---------------------------------------------------------------------------------------------
// probability_of_match, exact fp64 (reference :383-455 with :457-544 inlined), in two halves:
// match_prepare does the gates and the two Mahalanobis forms, match_finish the log / exp tail.
// ---------------------------------------------------------------------------------------------
// PK_MATCH_SKIP selects how a sole candidate is decided (all variants give identical results):
//   0  always evaluate the full likelihood;
//   1  prove positivity from the determinants' and forms' ranges (skips both logs and both exps);
//   2  evaluate both pdf exponents (logs overlap with the bearing chain), skip only the exps.
#ifndef PK_MATCH_SKIP
#define PK_MATCH_SKIP 2
#endif

struct MatchPre {
#if PK_MATCH_SKIP == 2
    double a2, a3;  // exponents of the position / colour pdf
#else
    double maha2, det2, maha3, det3;
#endif
    double pse;
    bool gated;
    // `sure`: the likelihood is PROVABLY a positive, finite, normal double, so a caller that only needs
    // the reference's match / no-match decision (`probability > 0.0`, :369-381) may skip match_finish.
    // Proof (variant 2): a2, a3 in (-700, 700) keep bp = e^a2 and cp = e^a3, and 500*bp, 500*cp, inside the
    // normal range; -700 < a2 + a3 < 690 keeps (500 bp)(500 cp) = 250000 e^(a2+a3) and its quotient by
    // 250000 normal and finite -- nothing underflows to zero or overflows whatever the last-bit
    // rounding of exp.  (Variant 1: det in (1e-100, 1e100) => |log det| < 230.3 and |maha| < 400 put both
    // exponents inside (-318, 314), same conclusion.)
    bool sure;
};

__device__ __forceinline__ MatchPre match_prepare(const Landmark& L, double px, double py, double pth, double beta,
                                                  double orr, double og, double ob, double dirx, double diry,
                                                  const pk_params& prm, unsigned& flags) {
    // Written without early exits: candidates that reach this point almost always pass the gates, and
    // one straight-line block lets the independent chains (bearing, position form, colour form) overlap.
    MatchPre m;
    // colour gate :425-427, :441
    const double dr = orr - L.r, dg = og - L.g, db = ob - L.b;
    const double cdist = dr * dr + dg * dg + db * db;
    bool gated = fabs(cdist) > prm.color_gate;
    // bearing gate :408-415, :433
    const double dx = L.x - px, dy = L.y - py;
    const double pse = pk_atan2(dy, dx);
    m.pse = pse;
    const double del = beta - (pse - pth);
    gated = gated || (fabs(del) > prm.bearing_gate);
    // prob_position_match :473-475 (robot-frame bearing used as if world frame, finding F4c)
    gated = gated || (fabs(pse - beta) > prm.position_gate);
    // closest_point :509-522
    const double t = dx * dirx + dy * diry;
    const double nx = (t < 0.0) ? px : px + dirx * t;
    const double ny = (t < 0.0) ? py : py + diry * t;
    const double ex = nx - L.x, ey = ny - L.y;
    // 2-D pdf, covariance symmetrised from the LOWER triangle (scipy eigh(lower=True)) :482-490
    const double a = L.sp[0], b10 = L.sp[2], d = L.sp[3];
    const double det2 = a * d - b10 * b10;
    const double maha2 = (d * ex * ex - 2.0 * b10 * ex * ey + a * ey * ey) / det2;
    // 3-D colour pdf :530-544, lower triangle
    const double A = L.sc[0], B = L.sc[3], C = L.sc[6], D = L.sc[4], E = L.sc[7], F = L.sc[8];
    const double c00 = D * F - E * E, c01 = C * E - B * F, c02 = B * E - C * D;
    const double c11 = A * F - C * C, c12 = B * C - A * E, c22 = A * D - B * B;
    const double det3 = A * c00 + B * c01 + C * c02;
    const double maha3 =
        (c00 * dr * dr + c11 * dg * dg + c22 * db * db + 2.0 * (c01 * dr * dg + c02 * dr * db + c12 * dg * db)) / det3;
    if (!gated && (!(det2 > 0.0) || !(det3 > 0.0))) flags |= PK_FLAG_SINGULAR_COV;
    m.gated = gated;
#if PK_MATCH_SKIP == 2
    m.a2 = -0.5 * (2.0 * kLog2Pi + log(det2) + maha2);
    m.a3 = -0.5 * (3.0 * kLog2Pi + log(det3) + maha3);
    const double a23 = m.a2 + m.a3;
    m.sure = !gated && (m.a2 > -700.0) && (m.a2 < 700.0) && (m.a3 > -700.0) && (m.a3 < 700.0) && (a23 > -700.0) &&
             (a23 < 690.0);
#else
    m.det2 = det2;
    m.maha2 = maha2;
    m.det3 = det3;
    m.maha3 = maha3;
    m.sure = (PK_MATCH_SKIP == 1) && !gated && (det2 > 1e-100) && (det2 < 1e100) && (det3 > 1e-100) && (det3 < 1e100) &&
             (fabs(maha2) < 400.0) && (fabs(maha3) < 400.0);
#endif
    return m;
}

__device__ __forceinline__ double match_finish(const MatchPre& m) {
#if PK_MATCH_SKIP == 2
    const double bp = pk_exp(m.a2);
    const double cp = pk_exp(m.a3);
#else
    const double bp = pk_exp(-0.5 * (2.0 * kLog2Pi + log(m.det2) + m.maha2));
    const double cp = pk_exp(-0.5 * (3.0 * kLog2Pi + log(m.det3) + m.maha3));
#endif
    // :439, :446, :455
    const double Lk = (500.0 * bp) * (500.0 * cp) / 250000.0;
    return m.gated ? 0.0 : Lk;
}

__device__ __forceinline__ double match_likelihood(const Landmark& L, double px, double py, double pth, double beta,
                                                   double orr, double og, double ob, double dirx, double diry,
                                                   const pk_params& prm, unsigned& flags, double& pse_out) {
    const MatchPre m = match_prepare(L, px, py, pth, beta, orr, og, ob, dirx, diry, prm, flags);
    pse_out = m.pse;
    return match_finish(m);
}

// ---------------------------------------------------------------------------------------------
// One EKF update of landmark j with blob k, block form of reference :98-124.
// Returns the weight factor; writes the landmark back unless it is immutable.
// ---------------------------------------------------------------------------------------------
// wrap an angle difference to (-pi, pi] (PK_MODEL_TEXTBOOK only; the reference never wraps, finding F4e)
__device__ __forceinline__ double pk_wrap_pi(double a) {
    const double two_pi = 6.283185307179586, pi = 3.141592653589793;
    a = a - two_pi * rint(a / two_pi);
    return (a <= -pi) ? a + two_pi : a;
}
__device__ __forceinline__ float pk_wrap_pi(float a) {
    const float two_pi = 6.283185307179586f, pi = 3.141592653589793f;
    a = a - two_pi * rintf(a * 0.15915494309189535f);
    return (a <= -pi) ? a + two_pi : a;
}

// `pth`: the particle's heading, used by PK_MODEL_TEXTBOOK only.
__device__ __forceinline__ double ekf_update_lm(Landmark& L, double px, double py, double beta, double orr, double og,
                                                double ob, const pk_params& prm, int& id_out, unsigned& flags,
                                                int& promoted, bool& changed_out, bool have_zb = false,
                                                double zb_in = 0.0, double pth = 0.0, bool have_log_nm = false,
                                                double log_nm = 0.0) {
    id_out = L.id;
    const double qt = prm.qt_diag;
    // measurement_jacobian :785-797 (sign and order as written, finding F4b)
    double dx = L.x - px, dy = L.y - py;
    double q = dx * dx + dy * dy;
    double hx = (q == 0.0) ? 0.0 : dy / q;
    double hy = (q == 0.0) ? 0.0 : dx / q;
    // generate_measurement :871 -- world-frame bearing, no heading subtraction (finding F4a)
    // (the association step evaluates the same atan2(fy - sy, fx - sx), :408/:473; reuse it)
    double zb = have_zb ? zb_in : pk_atan2(dy, dx);
    const bool textbook = (prm.model & PK_MODEL_TEXTBOOK) != 0;
    if (textbook) {  // d bearing / d(landmark x) = -dy/q, and the robot-frame prediction
        hx = -hx;
        zb = zb - pth;
    }
    double a = L.sp[0], b = L.sp[1], c = L.sp[2], d = L.sp[3];
    // measurement_covariance :817-819   Q = H Sigma H^T + Qt = diag(s) (+) Sc
    double t0 = hx * a + hy * c, t1 = hx * b + hy * d;
    double s = t0 * hx + t1 * hy + qt;
    double S00 = L.sc[0] + qt, S01 = L.sc[1], S02 = L.sc[2];
    double S10 = L.sc[3], S11 = L.sc[4] + qt, S12 = L.sc[5];
    double S20 = L.sc[6], S21 = L.sc[7], S22 = L.sc[8] + qt;
    // inverse(Q) :102 -- 1x1 block and general 3x3 block
    double inv_s = 1.0 / s;
    double C00 = S11 * S22 - S12 * S21, C01 = S12 * S20 - S10 * S22, C02 = S10 * S21 - S11 * S20;
    double detS = S00 * C00 + S01 * C01 + S02 * C02;
    if (!(detS != 0.0)) flags |= PK_FLAG_SINGULAR_COV;
    double idet = 1.0 / detS;
    double I00 = C00 * idet, I01 = (S02 * S21 - S01 * S22) * idet, I02 = (S01 * S12 - S02 * S11) * idet;
    double I10 = C01 * idet, I11 = (S00 * S22 - S02 * S20) * idet, I12 = (S02 * S10 - S00 * S12) * idet;
    double I20 = C02 * idet, I21 = (S01 * S20 - S00 * S21) * idet, I22 = (S00 * S11 - S01 * S10) * idet;
    // innovation :911 / :846 -- no angle wrapping (finding F4e)
    double d0 = beta - zb, d1 = orr - L.r, d2 = og - L.g, d3 = ob - L.b;
    if (textbook) d0 = pk_wrap_pi(d0);
    // importance_factor :844-849 with the PRE-update Q and z-hat; Frobenius norm of Q (F4d)
    double fro = sqrt(s * s + S00 * S00 + S01 * S01 + S02 * S02 + S10 * S10 + S11 * S11 + S12 * S12 + S20 * S20 +
                      S21 * S21 + S22 * S22);
    double y1 = d1 * I00 + d2 * I10 + d3 * I20;  // (delz^T Qinv) colour part
    double y2 = d1 * I01 + d2 * I11 + d3 * I21;
    double y3 = d1 * I02 + d2 * I12 + d3 * I22;
    double maha = d0 * inv_s * d0 + y1 * d1 + y2 * d2 + y3 * d3;
    const bool log_w = (prm.model & PK_MODEL_LOG_WEIGHTS) != 0;
    double factor = log_w ? -0.5 * (log(2.0 * 3.141592653589793 * fro) + maha)
                          : (1.0 / sqrt(2.0 * 3.141592653589793 * fro)) * pk_exp(-0.5 * maha);

    bool changed = false;
    if (!(L.meta & PK_META_IMMUTABLE)) {
        // kalman_gain :833   K = Sigma H^T Qinv
        double kp0 = (a * hx + b * hy) * inv_s, kp1 = (c * hx + d * hy) * inv_s;
        const double* sc = L.sc;
        double K00 = sc[0] * I00 + sc[1] * I10 + sc[2] * I20, K01 = sc[0] * I01 + sc[1] * I11 + sc[2] * I21,
               K02 = sc[0] * I02 + sc[1] * I12 + sc[2] * I22;
        double K10 = sc[3] * I00 + sc[4] * I10 + sc[5] * I20, K11 = sc[3] * I01 + sc[4] * I11 + sc[5] * I21,
               K12 = sc[3] * I02 + sc[4] * I12 + sc[5] * I22;
        double K20 = sc[6] * I00 + sc[7] * I10 + sc[8] * I20, K21 = sc[6] * I01 + sc[7] * I11 + sc[8] * I21,
               K22 = sc[6] * I02 + sc[7] * I12 + sc[8] * I22;
        // update_mean :909-914
        L.x += kp0 * d0;
        L.y += kp1 * d0;
        L.r += K00 * d1 + K01 * d2 + K02 * d3;
        L.g += K10 * d1 + K11 * d2 + K12 * d3;
        L.b += K20 * d1 + K21 * d2 + K22 * d3;
        // update_covar :926-930   Sigma <- (I - K H) Sigma
        double m00 = 1.0 - kp0 * hx, m01 = -(kp0 * hy), m10 = -(kp1 * hx), m11 = 1.0 - kp1 * hy;
        L.sp[0] = m00 * a + m01 * c;
        L.sp[1] = m00 * b + m01 * d;
        L.sp[2] = m10 * a + m11 * c;
        L.sp[3] = m10 * b + m11 * d;
        double A00 = 1.0 - K00, A01 = -K01, A02 = -K02;
        double A10 = -K10, A11 = 1.0 - K11, A12 = -K12;
        double A20 = -K20, A21 = -K21, A22 = 1.0 - K22;
        double n0 = A00 * sc[0] + A01 * sc[3] + A02 * sc[6], n1 = A00 * sc[1] + A01 * sc[4] + A02 * sc[7],
               n2 = A00 * sc[2] + A01 * sc[5] + A02 * sc[8];
        double n3 = A10 * sc[0] + A11 * sc[3] + A12 * sc[6], n4 = A10 * sc[1] + A11 * sc[4] + A12 * sc[7],
               n5 = A10 * sc[2] + A11 * sc[5] + A12 * sc[8];
        double n6 = A20 * sc[0] + A21 * sc[3] + A22 * sc[6], n7 = A20 * sc[1] + A21 * sc[4] + A22 * sc[7],
               n8 = A20 * sc[2] + A21 * sc[5] + A22 * sc[8];
        L.sc[0] = n0; L.sc[1] = n1; L.sc[2] = n2;
        L.sc[3] = n3; L.sc[4] = n4; L.sc[5] = n5;
        L.sc[6] = n6; L.sc[7] = n7; L.sc[8] = n8;
        int cnt = (L.meta & PK_META_COUNT_MASK) + 2;  // :914 and :930, +1 each
        if (cnt > PK_META_COUNT_MASK) cnt = PK_META_COUNT_MASK;
        L.meta = (L.meta & ~PK_META_COUNT_MASK) | cnt;
        changed = true;
    }
    if (id_out < 0) {
        // potential feature :109-118: weight as if unseen; promote when update_count > 5
        // (K2 passes log(no_match_weight) in: an fp64 log set up inside its main loop costs every group)
        factor = log_w ? (have_log_nm ? log_nm : log(prm.no_match_weight)) : prm.no_match_weight;
        if (L.meta & PK_META_POTENTIAL) {
            if ((L.meta & PK_META_COUNT_MASK) > prm.promote_count) {
                L.meta &= ~PK_META_POTENTIAL;
                L.id = -L.id;
                promoted += 1;
                changed = true;
            }
        }
    }
    changed_out = changed;
    return factor;
}

// ---------------------------------------------------------------------------------------------
// fp32 landmark algebra (PK_DTYPE_ARITH_F32): the same formulas on the fp32 record, symmetric blocks as
// stored (lower triangles).  Differences pose - landmark are formed in fp64 (poses are fp64 and may be far from
// the origin) and then rounded; the importance factor's exp and everything downstream of it is fp64.  No exp
// is needed for the association: the reference's decision `probability > 0.0` is an fp64 underflow test
// (finding F3), equivalent to thresholds on the two pdf exponents, and ranking several candidates by
// bp*cp is ranking by a2 + a3.
// ---------------------------------------------------------------------------------------------
// Branch-free atan2 in fp32: the reduction of pk_atan2 (one division, |t| <= tan(pi/8)) with the degree-9 odd
// polynomial of Cephes' atanf (error < 2e-7 rad); fast division (2 ulp).
__device__ __forceinline__ float pk_atan2f(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const bool steep = ay > ax;
    const float mx = steep ? ay : ax, mn = steep ? ax : ay;
    const bool big = mn > 0.41421356f * mx;
    const float num = big ? mn - mx : mn;
    const float den = big ? mn + mx : mx;
    float t = __fdividef(num, den);
    if (den == 0.0f) t = 0.0f;
    const float z = t * t;
    const float p = ((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f;
    float r = fmaf(t * z, p, t);
    if (big) r += 0.78539816339744831f;
    if (steep) r = 1.57079632679489662f - r;
    if (signbit(x)) r = 3.14159265358979324f - r;
    return copysignf(r, y);
}

// exp(x) for an fp32 argument with an fp64 RESULT (the importance factor spans hundreds of decades, finding F3, but
// its argument -maha/2 comes out of fp32 algebra): n = rint(x log2 e), 2^f by the hardware ex2 on the compensated
// remainder f = x log2 e - n (|f| <= 1/2, abs. error ~3e-8), scaled by 2^n in two exact steps so that the last
// multiply rounds once into the subnormal range (or to 0).  Relative error ~3e-7, fifteen instructions instead of
// the fifty of the fp64 polynomial.
__device__ __forceinline__ double pk_exp_f2d(float x) {
    const float t = x * 1.44269502162933349609375f;
    const float n = rintf(t);
    float f = fmaf(x, 1.44269502162933349609375f, -n);
    f = fmaf(x, 1.925963033500011e-8f, f);
    const float m = exp2f(f);
    int ni = (int)fminf(fmaxf(n, -1100.0f), 1100.0f);
    const int n1 = ni >> 1, n2 = ni - n1;
    const double s1 = __hiloint2double((n1 + 1023) << 20, 0);
    const double s2 = __hiloint2double((n2 + 1023) << 20, 0);
    return ((double)m * s1) * s2;
}

struct MatchPreF {
    float a2, a3, pse;
    bool gated;
    bool sure;  // == the likelihood is positive in fp64: a2, a3 and a2 + a3 above log(2^-1075) = -745.13
};

constexpr float kLog2PiF = 1.8378770664093453f;
constexpr float kUnderflowF = -745.13f;

__device__ __forceinline__ MatchPreF match_prepare(const LandmarkF& L, double px, double py, double pth, float beta, float orr,
                                                   float og, float ob, float dirx, float diry, const pk_params& prm,
                                                   unsigned& flags) {
    MatchPreF m;
    const float dr = orr - L.r, dg = og - L.g, db = ob - L.b;
    const float cdist = dr * dr + dg * dg + db * db;
    bool gated = fabsf(cdist) > (float)prm.color_gate;                                  // :441
    const float dx = (float)((double)L.x - px), dy = (float)((double)L.y - py);
    const float pse = pk_atan2f(dy, dx);                                                // :408 / :473
    m.pse = pse;
    const float del = beta - (pse - (float)pth);
    gated = gated || (fabsf(del) > (float)prm.bearing_gate);                            // :433
    gated = gated || (fabsf(pse - beta) > (float)prm.position_gate);                    // :474
    const float t = dx * dirx + dy * diry;                                              // closest_point :509-522
    const float ex = (t < 0.0f) ? -dx : dirx * t - dx;                                  // near - landmark
    const float ey = (t < 0.0f) ? -dy : diry * t - dy;
    const float a = L.sp[0], b10 = L.sp[1], d = L.sp[2];
    const float det2 = a * d - b10 * b10;
    const float maha2 = __fdividef(d * ex * ex - 2.0f * b10 * ex * ey + a * ey * ey, det2);
    const float A = L.sc[0], B = L.sc[1], C = L.sc[3], D = L.sc[2], E = L.sc[4], F = L.sc[5];
    const float c00 = D * F - E * E, c01 = C * E - B * F, c02 = B * E - C * D;
    const float c11 = A * F - C * C, c12 = B * C - A * E, c22 = A * D - B * B;
    const float det3 = A * c00 + B * c01 + C * c02;
    const float maha3 = __fdividef(
        c00 * dr * dr + c11 * dg * dg + c22 * db * db + 2.0f * (c01 * dr * dg + c02 * dr * db + c12 * dg * db), det3);
    if (!gated && (!(det2 > 0.0f) || !(det3 > 0.0f))) flags |= PK_FLAG_SINGULAR_COV;
    m.gated = gated;
    // the exponents only feed decisions and rankings: the hardware log2 (abs. error ~1e-6 in the log) is enough
    m.a2 = -0.5f * (2.0f * kLog2PiF + __logf(det2) + maha2);
    m.a3 = -0.5f * (3.0f * kLog2PiF + __logf(det3) + maha3);
    m.sure = !gated && (m.a2 > kUnderflowF) && (m.a3 > kUnderflowF) && (m.a2 + m.a3 > kUnderflowF);
    return m;
}

// rank value: 0 = no match, otherwise increasing with the likelihood (log domain, shifted positive)
__device__ __forceinline__ float match_finish(const MatchPreF& m) {
    return m.sure ? (m.a2 + m.a3) + 2048.0f : 0.0f;
}

__device__ __forceinline__ float match_likelihood(const LandmarkF& L, double px, double py, double pth, float beta, float orr,
                                                   float og, float ob, float dirx, float diry, const pk_params& prm,
                                                   unsigned& flags, float& pse_out) {
    const MatchPreF m = match_prepare(L, px, py, pth, beta, orr, og, ob, dirx, diry, prm, flags);
    pse_out = m.pse;
    return match_finish(m);
}

__device__ __forceinline__ double ekf_update_lm(LandmarkF& L, double px, double py, float beta, float orr, float og, float ob,
                                                const pk_params& prm, int& id_out, unsigned& flags, int& promoted,
                                                bool& changed_out, bool have_zb = false, float zb_in = 0.0f, double pth = 0.0,
                                                bool have_log_nm = false, double log_nm = 0.0) {
    id_out = L.id;
    const float qt = (float)prm.qt_diag;
    const float dx = (float)((double)L.x - px), dy = (float)((double)L.y - py);
    const float q = dx * dx + dy * dy;                                                   // :785
    const float inv_q = __fdividef(1.0f, q);
    float hx = (q == 0.0f) ? 0.0f : dy * inv_q;                                          // :788-797 (as written, F4b)
    const float hy = (q == 0.0f) ? 0.0f : dx * inv_q;
    float zb = zb_in;                                                                    // :871 (world frame, F4a)
    if (!have_zb) zb = pk_atan2f(dy, dx);
    const bool textbook = (prm.model & PK_MODEL_TEXTBOOK) != 0;
    if (textbook) {
        hx = -hx;
        zb = zb - (float)pth;
    }
    const float a = L.sp[0], b = L.sp[1], d = L.sp[2];                                   // S00, S10 (= S01), S11
    const float t0 = hx * a + hy * b, t1 = hx * b + hy * d;
    const float s = t0 * hx + t1 * hy + qt;                                              // :817-819
    const float S00 = L.sc[0] + qt, S10 = L.sc[1], S11 = L.sc[2] + qt, S20 = L.sc[3], S21 = L.sc[4], S22 = L.sc[5] + qt;
    const float inv_s = __fdividef(1.0f, s);
    // symmetric 3x3 inverse
    const float C00 = S11 * S22 - S21 * S21, C01 = S21 * S20 - S10 * S22, C02 = S10 * S21 - S11 * S20;
    const float detS = S00 * C00 + S10 * C01 + S20 * C02;
    if (!(detS != 0.0f)) flags |= PK_FLAG_SINGULAR_COV;
    const float idet = __fdividef(1.0f, detS);
    const float I00 = C00 * idet, I10 = C01 * idet, I20 = C02 * idet;
    const float I11 = (S00 * S22 - S20 * S20) * idet, I21 = (S20 * S10 - S00 * S21) * idet, I22 = (S00 * S11 - S10 * S10) * idet;
    float d0 = beta - zb;                                                                // :911 / :846, no wrapping (F4e)
    if (textbook) d0 = pk_wrap_pi(d0);
    const float d1 = orr - L.r, d2 = og - L.g, d3 = ob - L.b;
    const float fro = sqrtf(s * s + S00 * S00 + S11 * S11 + S22 * S22 + 2.0f * (S10 * S10 + S20 * S20 + S21 * S21));
    const float y1 = d1 * I00 + d2 * I10 + d3 * I20;
    const float y2 = d1 * I10 + d2 * I11 + d3 * I21;
    const float y3 = d1 * I20 + d2 * I21 + d3 * I22;
    const float maha = d0 * inv_s * d0 + y1 * d1 + y2 * d2 + y3 * d3;
    // importance_factor :844-849: the exp and the product that follows stay fp64 (weights span hundreds of decades)
    const bool log_w = (prm.model & PK_MODEL_LOG_WEIGHTS) != 0;
    double factor = log_w ? (double)(-0.5f * (__logf(2.0f * 3.14159265358979f * fro) + maha))
                          : (double)rsqrtf(2.0f * 3.14159265358979f * fro) * pk_exp_f2d(-0.5f * maha);

    bool changed = false;
    if (!(L.meta & PK_META_IMMUTABLE)) {
        // kalman_gain :833  K = Sigma H^T Qinv.  Position block: kp = Sigma_p h / s.  Colour block: Kc = Sigma_c Sc^-1,
        // symmetric because Sc = Sigma_c + qt I commutes with Sigma_c -- six entries instead of nine.
        const float kp0 = t0 * inv_s, kp1 = t1 * inv_s;
        const float c00 = L.sc[0], c10 = L.sc[1], c11 = L.sc[2], c20 = L.sc[3], c21 = L.sc[4], c22 = L.sc[5];
        const float K00 = c00 * I00 + c10 * I10 + c20 * I20;
        const float K10 = c10 * I00 + c11 * I10 + c21 * I20;
        const float K11 = c10 * I10 + c11 * I11 + c21 * I21;
        const float K20 = c20 * I00 + c21 * I10 + c22 * I20;
        const float K21 = c20 * I10 + c21 * I11 + c22 * I21;
        const float K22 = c20 * I20 + c21 * I21 + c22 * I22;
        L.x += kp0 * d0;                                                                  // :909-914
        L.y += kp1 * d0;
        L.r += K00 * d1 + K10 * d2 + K20 * d3;
        L.g += K10 * d1 + K11 * d2 + K21 * d3;
        L.b += K20 * d1 + K21 * d2 + K22 * d3;
        // update_covar :926-930  Sigma <- (I - K H) Sigma, as algebra on the stored (lower-triangle) blocks:
        // position  (I - kp h^T) Sigma_p = Sigma_p - kp (Sigma_p h)^T;  colour  (I - Kc) Sigma_c = qt Sc^-1 Sigma_c = qt Kc
        L.sp[0] = a - kp0 * t0;
        L.sp[1] = b - kp1 * t0;
        L.sp[2] = d - kp1 * t1;
        L.sc[0] = qt * K00;
        L.sc[1] = qt * K10;
        L.sc[2] = qt * K11;
        L.sc[3] = qt * K20;
        L.sc[4] = qt * K21;
        L.sc[5] = qt * K22;
        int cnt = (L.meta & PK_META_COUNT_MASK) + 2;
        if (cnt > PK_META_COUNT_MASK) cnt = PK_META_COUNT_MASK;
        L.meta = (L.meta & ~PK_META_COUNT_MASK) | cnt;
        changed = true;
    }
    if (id_out < 0) {
        // (K2 passes log(no_match_weight) in: an fp64 log set up inside its main loop costs every group)
        factor = log_w ? (have_log_nm ? log_nm : log(prm.no_match_weight)) : prm.no_match_weight;
        if (L.meta & PK_META_POTENTIAL) {
            if ((L.meta & PK_META_COUNT_MASK) > prm.promote_count) {
                L.meta &= ~PK_META_POTENTIAL;
                L.id = -L.id;
                promoted += 1;
                changed = true;
            }
        }
    }
    changed_out = changed;
    return factor;
}

template <typename T>
__device__ __forceinline__ double ekf_update(unsigned char* block, int capacity, int j, double px, double py,
                                             double beta, double orr, double og, double ob, const pk_params& prm,
                                             int& id_out, unsigned& flags, int& promoted, double pth = 0.0) {
    Landmark L;
    load_landmark<T>(block, capacity, j, L);
    bool changed = false;
    const double factor = ekf_update_lm(L, px, py, beta, orr, og, ob, prm, id_out, flags, promoted, changed, false, 0.0, pth);
    if (changed) store_landmark<T>(block, capacity, j, L);
    return factor;
}

}  // namespace pk
