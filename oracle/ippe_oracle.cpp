// oracle/ippe_oracle.cpp -- TEST INFRASTRUCTURE (CPU checker), not product code.
//
// Restatement of the marker pose step of the reference:
//   aruco::Marker::calculateExtrinsics        Thirdparty/aruco/aruco/marker.cpp:322-343  (get3DPoints :358-369)
//   aruco::solvePnP -> IPPE::PoseSolver::solveGeneric   Thirdparty/aruco/aruco/ippe.cpp:72-169
//   solveGeneric(object, normalized)           ippe.cpp:171-223   makeCanonicalObjectPoints  ippe.cpp:647-746
//   HomographyHO::homographyHO                 ippe.cpp:809-1033  solveCanonicalForm         ippe.cpp:225-264
//   computeRotations / rotateVec2ZAxis         ippe.cpp:485-590, 1036-1080
//   computeTranslation                         ippe.cpp:395-483   rot2vec                    ippe.cpp:365-393
//   sortPosesByReprojError / evalReprojError   ippe.cpp:748-807
// and of the OpenCV calls inside it (OpenCV is an un-vendored dependency of the reference; semantics of 4.13, the
// only version that can be executed here): cv::undistortPoints (5 fixed-point iterations, result stored as float because
// the input is vector<Point2f>), cv::eigen on the symmetric 3x3 (Jacobi), cv::Rodrigues (vector -> matrix) and
// cv::projectPoints (result stored as float because the object points are Point3f).
// Pinned by the reference itself: Thirdparty/aruco/aruco/ippe.cpp compiles UNMODIFIED on oracle/ippeshim into oracle/_ref/libref_ippe.so
// (oracle/ref_ippe_wrap.cpp does what Marker::calculateExtrinsics does); this restatement equals it on every marker tried
// (tests/test_oracle_ippe_vs_ref.py, golden tests/golden/ippe_ref.npz for boxes without the reference: rvec / tvec of both solutions <= 1e-7,
// errors <= 1e-6 relative).  Also checked against cv2.solvePnPGeneric(SOLVEPNP_IPPE) (the same author's algorithm inside OpenCV) and
// against ground-truth poses of projected squares (tests/test_oracle_ippe.py).
//
// Arithmetic: double, except where the reference rounds to float (normalized points, projected points, error sums,
// the returned Rvec/Tvec).  Type quirks kept: with float object points makeCanonicalObjectPoints never leaves the
// z-plane branch (the |z| test sits in the double branch only, ippe.cpp:675-679).
#include <math.h>
#include <float.h>
#include <string.h>
#include "oracle.h"

namespace {

struct Cam { double fx, fy, cx, cy; double k[5]; };      // k1 k2 p1 p2 k3

// cv::undistortPoints without R/P: pixel -> normalized coordinates, stored as float (dst has the type of src)
void undistort_point(const Cam& c, float u, float v, float& xo, float& yo) {
    double x = ((double)u - c.cx) * (1.0 / c.fx), y = ((double)v - c.cy) * (1.0 / c.fy);
    const double x0 = x, y0 = y;
    const bool has_dist = c.k[0] != 0 || c.k[1] != 0 || c.k[2] != 0 || c.k[3] != 0 || c.k[4] != 0;
    if (has_dist) {
        for (int j = 0; j < 5; j++) {
            const double r2 = x * x + y * y;
            const double icdist = 1.0 / (1 + ((c.k[4] * r2 + c.k[1]) * r2 + c.k[0]) * r2);
            if (icdist < 0) { x = x0; y = y0; break; }
            const double dx = 2 * c.k[2] * x * y + c.k[3] * (r2 + 2 * x * x);
            const double dy = c.k[2] * (r2 + 2 * y * y) + 2 * c.k[3] * x * y;
            x = (x0 - dx) * icdist;
            y = (y0 - dy) * icdist;
        }
    }
    xo = (float)x; yo = (float)y;
}

// symmetric 3x3 eigen decomposition, cyclic Jacobi with the largest off-diagonal element as pivot; eigenvalues
// descending, eigenvectors as ROWS of V (the layout of cv::eigen)
void eigen_sym3(const double Ain[9], double W[3], double V[9]) {
    double A[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { A[i][j] = Ain[3 * i + j]; V[3 * i + j] = i == j ? 1.0 : 0.0; }
    for (int it = 0; it < 270; it++) {
        int k = 0, l = 1;
        double mv = fabs(A[0][1]);
        if (fabs(A[0][2]) > mv) { mv = fabs(A[0][2]); k = 0; l = 2; }
        if (fabs(A[1][2]) > mv) { mv = fabs(A[1][2]); k = 1; l = 2; }
        const double p = A[k][l];
        if (fabs(p) <= DBL_EPSILON) break;
        const double y = (A[l][l] - A[k][k]) * 0.5;
        double t = fabs(y) + hypot(p, y);
        double s = hypot(p, t);
        const double c = t / s;
        s = p / s; t = (p / t) * p;
        if (y < 0) { s = -s; t = -t; }
        A[k][l] = 0;
        A[k][k] -= t; A[l][l] += t;
#define ROT(v0, v1) { const double a0 = v0, b0 = v1; v0 = a0 * c - b0 * s; v1 = a0 * s + b0 * c; }
        for (int i = 0; i < k; i++) ROT(A[i][k], A[i][l]);
        for (int i = k + 1; i < l; i++) ROT(A[k][i], A[i][l]);
        for (int i = l + 1; i < 3; i++) ROT(A[k][i], A[l][i]);
        for (int i = 0; i < 3; i++) ROT(V[3 * k + i], V[3 * l + i]);
#undef ROT
    }
    for (int i = 0; i < 3; i++) W[i] = A[i][i];
    for (int k = 0; k < 2; k++) {
        int m = k;
        for (int i = k + 1; i < 3; i++) if (W[m] < W[i]) m = i;
        if (m != k) { double tw = W[m]; W[m] = W[k]; W[k] = tw; for (int i = 0; i < 3; i++) { double tv = V[3 * m + i]; V[3 * m + i] = V[3 * k + i]; V[3 * k + i] = tv; } }
    }
}

// HomographyHO::normalizeDataIsotropic (ippe.cpp:809-910): D[2][n], T, Ti
void normalize_iso(const double* xs, const double* ys, int n, double* D0, double* D1, double T[9], double Ti[9]) {
    double xm = 0, ym = 0;
    for (int i = 0; i < n; i++) { xm += xs[i]; ym += ys[i]; }
    xm /= (double)n; ym /= (double)n;
    double kappa = 0;
    for (int i = 0; i < n; i++) { D0[i] = xs[i] - xm; D1[i] = ys[i] - ym; kappa += D0[i] * D0[i] + D1[i] * D1[i]; }
    const double beta = sqrt(2 * n / kappa);
    for (int i = 0; i < n; i++) { D0[i] *= beta; D1[i] *= beta; }
    memset(T, 0, 72); memset(Ti, 0, 72);
    T[0] = 1.0 / beta; T[4] = 1.0 / beta; T[2] = xm; T[5] = ym; T[8] = 1;
    Ti[0] = beta; Ti[4] = beta; Ti[2] = -beta * xm; Ti[5] = -beta * ym; Ti[8] = 1;
}

void mat3_mul(const double A[9], const double B[9], double C[9]) {
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += A[3 * i + k] * B[3 * k + j]; C[3 * i + j] = s; }
}

// HomographyHO::homographyHO for n = 4 (ippe.cpp:912-1033)
void homography_ho(const double* ax, const double* ay, const double* bx, const double* by, double H[9]) {
    const int n = 4;
    double A0[4], A1[4], B0[4], B1[4], TA[9], TAi[9], TB[9], TBi[9];
    normalize_iso(ax, ay, n, A0, A1, TA, TAi);
    normalize_iso(bx, by, n, B0, B1, TB, TBi);
    double C1[4], C2[4], C3[4], C4[4], mC1 = 0, mC2 = 0, mC3 = 0, mC4 = 0;
    for (int i = 0; i < n; i++) {
        C1[i] = -B0[i] * A0[i]; C2[i] = -B0[i] * A1[i]; C3[i] = -B1[i] * A0[i]; C4[i] = -B1[i] * A1[i];
        mC1 += C1[i]; mC2 += C2[i]; mC3 += C3[i]; mC4 += C4[i];
    }
    mC1 /= n; mC2 /= n; mC3 /= n; mC4 /= n;
    double Mx[4][3], My[4][3];
    for (int i = 0; i < n; i++) {
        Mx[i][0] = C1[i] - mC1; Mx[i][1] = C2[i] - mC2; Mx[i][2] = -B0[i];
        My[i][0] = C3[i] - mC3; My[i][1] = C4[i] - mC4; My[i][2] = -B1[i];
    }
    // DataA * DataA^T (2x2) and its inverse
    double g00 = 0, g01 = 0, g11 = 0;
    for (int i = 0; i < n; i++) { g00 += A0[i] * A0[i]; g01 += A0[i] * A1[i]; g11 += A1[i] * A1[i]; }
    const double dt = g00 * g11 - g01 * g01;
    const double i00 = g11 / dt, i01 = -g01 / dt, i10 = -g01 / dt, i11 = g00 / dt;
    double Pp0[4], Pp1[4];                        // Pp = inv * DataA (2 x n)
    for (int i = 0; i < n; i++) { Pp0[i] = i00 * A0[i] + i01 * A1[i]; Pp1[i] = i10 * A0[i] + i11 * A1[i]; }
    double Bx[2][3], By[2][3];                    // Pp * Mx, Pp * My
    for (int j = 0; j < 3; j++) {
        double s0 = 0, s1 = 0, t0 = 0, t1 = 0;
        for (int i = 0; i < n; i++) { s0 += Pp0[i] * Mx[i][j]; s1 += Pp1[i] * Mx[i][j]; t0 += Pp0[i] * My[i][j]; t1 += Pp1[i] * My[i][j]; }
        Bx[0][j] = s0; Bx[1][j] = s1; By[0][j] = t0; By[1][j] = t1;
    }
    double D[8][3];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < 3; j++) {
            const double ex = A0[i] * Bx[0][j] + A1[i] * Bx[1][j], ey = A0[i] * By[0][j] + A1[i] * By[1][j];     // DataA^T * Bx
            D[i][j] = Mx[i][j] - ex; D[i + n][j] = My[i][j] - ey;
        }
    double DDT[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 2 * n; k++) s += D[k][i] * D[k][j]; DDT[3 * i + j] = s; }
    double W[3], U[9];
    eigen_sym3(DDT, W, U);
    const double h7 = U[6], h8 = U[7], h9 = U[8];                  // eigenvector of the smallest eigenvalue
    const double h1 = -(Bx[0][0] * h7 + Bx[0][1] * h8 + Bx[0][2] * h9), h2 = -(Bx[1][0] * h7 + Bx[1][1] * h8 + Bx[1][2] * h9);
    const double h4 = -(By[0][0] * h7 + By[0][1] * h8 + By[0][2] * h9), h5 = -(By[1][0] * h7 + By[1][1] * h8 + By[1][2] * h9);
    const double h3 = -(mC1 * h7 + mC2 * h8), h6 = -(mC3 * h7 + mC4 * h8);
    const double Hn[9] = {h1, h2, h3, h4, h5, h6, h7, h8, h9};
    double T1[9];
    mat3_mul(TB, Hn, T1);
    mat3_mul(T1, TAi, H);
    const double s = H[8];
    for (int i = 0; i < 9; i++) H[i] = H[i] / s;
}

// IPPE::PoseSolver::rotateVec2ZAxis (ippe.cpp:1036-1080)
void rotate_vec2z(double ax, double ay, double az, double Ra[9]) {
    const double nrm = sqrt(ax * ax + ay * ay + az * az);
    ax /= nrm; ay /= nrm; az /= nrm;
    const double c = az;
    if (fabs(1.0 + c) < (double)FLT_EPSILON) {
        memset(Ra, 0, 72); Ra[0] = 1; Ra[4] = 1; Ra[8] = -1;
    } else {
        const double d = 1.0 / (1.0 + c), ax2 = ax * ax, ay2 = ay * ay, axay = ax * ay;
        Ra[0] = -ax2 * d + 1.0; Ra[1] = -axay * d; Ra[2] = -ax;
        Ra[3] = -axay * d; Ra[4] = -ay2 * d + 1.0; Ra[5] = -ay;
        Ra[6] = ax; Ra[7] = ay; Ra[8] = 1.0 - (ax2 + ay2) * d;
    }
}

// IPPE::PoseSolver::computeRotations (ippe.cpp:485-590)
void compute_rotations(double j00, double j01, double j10, double j11, double p, double q, double R1[9], double R2[9]) {
    double Rt[9], Rv[9];
    rotate_vec2z(p, q, 1.0, Rt);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Rv[3 * i + j] = Rt[3 * j + i];
    const double rv00 = Rv[0], rv01 = Rv[1], rv02 = Rv[2], rv10 = Rv[3], rv11 = Rv[4], rv12 = Rv[5], rv20 = Rv[6], rv21 = Rv[7], rv22 = Rv[8];
    const double b00 = rv00 - p * rv20, b01 = rv01 - p * rv21, b10 = rv10 - q * rv20, b11 = rv11 - q * rv21;
    const double dtinv = 1.0 / ((b00 * b11 - b01 * b10));
    const double binv00 = dtinv * b11, binv01 = -dtinv * b01, binv10 = -dtinv * b10, binv11 = dtinv * b00;
    const double a00 = binv00 * j00 + binv01 * j10, a01 = binv00 * j01 + binv01 * j11;
    const double a10 = binv10 * j00 + binv11 * j10, a11 = binv10 * j01 + binv11 * j11;
    const double ata00 = a00 * a00 + a01 * a01, ata01 = a00 * a10 + a01 * a11, ata11 = a10 * a10 + a11 * a11;
    const double gamma = sqrt(0.5 * (ata00 + ata11 + sqrt((ata00 - ata11) * (ata00 - ata11) + 4.0 * ata01 * ata01)));
    const double rt00 = a00 / gamma, rt01 = a01 / gamma, rt10 = a10 / gamma, rt11 = a11 / gamma;
    const double b0 = sqrt(-rt00 * rt00 - rt10 * rt10 + 1);
    double b1 = sqrt(-rt01 * rt01 - rt11 * rt11 + 1);
    const double sp = (-rt00 * rt01 - rt10 * rt11);
    if (sp < 0) b1 = -b1;
    const double c0 = b1 * rt10 - b0 * rt11, c1 = b0 * rt01 - b1 * rt00, c2 = rt00 * rt11 - rt01 * rt10;
    R1[0] = rt00 * rv00 + rt10 * rv01 + b0 * rv02; R1[1] = rt01 * rv00 + rt11 * rv01 + b1 * rv02; R1[2] = c0 * rv00 + c1 * rv01 + c2 * rv02;
    R1[3] = rt00 * rv10 + rt10 * rv11 + b0 * rv12; R1[4] = rt01 * rv10 + rt11 * rv11 + b1 * rv12; R1[5] = c0 * rv10 + c1 * rv11 + c2 * rv12;
    R1[6] = rt00 * rv20 + rt10 * rv21 + b0 * rv22; R1[7] = rt01 * rv20 + rt11 * rv21 + b1 * rv22; R1[8] = c0 * rv20 + c1 * rv21 + c2 * rv22;
    const double e0 = b0 * rt11 - b1 * rt10, e1 = b1 * rt00 - b0 * rt01;
    R2[0] = rt00 * rv00 + rt10 * rv01 + (-b0) * rv02; R2[1] = rt01 * rv00 + rt11 * rv01 + (-b1) * rv02; R2[2] = e0 * rv00 + e1 * rv01 + c2 * rv02;
    R2[3] = rt00 * rv10 + rt10 * rv11 + (-b0) * rv12; R2[4] = rt01 * rv10 + rt11 * rv11 + (-b1) * rv12; R2[5] = e0 * rv10 + e1 * rv11 + c2 * rv12;
    R2[6] = rt00 * rv20 + rt10 * rv21 + (-b0) * rv22; R2[7] = rt01 * rv20 + rt11 * rv21 + (-b1) * rv22; R2[8] = e0 * rv20 + e1 * rv21 + c2 * rv22;
}

// IPPE::PoseSolver::computeTranslation (ippe.cpp:395-483)
void compute_translation(const double* ox, const double* oy, const double* ix, const double* iy, int n, const double R[9], double t[3]) {
    const double ATA00 = n, ATA11 = n;
    double ATA02 = 0, ATA12 = 0, ATA20 = 0, ATA21 = 0, ATA22 = 0, ATb0 = 0, ATb1 = 0, ATb2 = 0;
    for (int i = 0; i < n; i++) {
        const double rx = R[0] * ox[i] + R[1] * oy[i], ry = R[3] * ox[i] + R[4] * oy[i], rz = R[6] * ox[i] + R[7] * oy[i];
        const double a2 = -ix[i], b2 = -iy[i];
        ATA02 = ATA02 + a2; ATA12 = ATA12 + b2; ATA20 = ATA20 + a2; ATA21 = ATA21 + b2; ATA22 = ATA22 + a2 * a2 + b2 * b2;
        const double bx = -a2 * rz - rx, by = -b2 * rz - ry;
        ATb0 = ATb0 + bx; ATb1 = ATb1 + by; ATb2 = ATb2 + a2 * bx + b2 * by;
    }
    const double detAInv = 1.0 / (ATA00 * ATA11 * ATA22 - ATA00 * ATA12 * ATA21 - ATA02 * ATA11 * ATA20);
    const double S00 = ATA11 * ATA22 - ATA12 * ATA21, S01 = ATA02 * ATA21, S02 = -ATA02 * ATA11;
    const double S10 = ATA12 * ATA20, S11 = ATA00 * ATA22 - ATA02 * ATA20, S12 = -ATA00 * ATA12;
    const double S20 = -ATA11 * ATA20, S21 = -ATA00 * ATA21, S22 = ATA00 * ATA11;
    t[0] = detAInv * (S00 * ATb0 + S01 * ATb1 + S02 * ATb2);
    t[1] = detAInv * (S10 * ATb0 + S11 * ATb1 + S12 * ATb2);
    t[2] = detAInv * (S20 * ATb0 + S21 * ATb1 + S22 * ATb2);
}

// IPPE::PoseSolver::rot2vec (ippe.cpp:365-393)
void rot2vec(const double R[9], double r[3]) {
    const double trace = R[0] + R[4] + R[8];
    const double w_norm = acos((trace - 1.0) / 2.0);
    const double eps = (double)FLT_EPSILON;
    const double d = 1 / (2 * sin(w_norm)) * w_norm;
    if (w_norm < eps) { r[0] = r[1] = r[2] = 0; return; }
    r[0] = d * (R[7] - R[5]); r[1] = d * (R[2] - R[6]); r[2] = d * (R[3] - R[1]);
}

// cv::Rodrigues, vector -> matrix (double)
void rodrigues(const double r[3], double R[9]) {
    const double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (theta < DBL_EPSILON) { memset(R, 0, 72); R[0] = R[4] = R[8] = 1; return; }
    const double c = cos(theta), s = sin(theta), c1 = 1. - c, it = 1. / theta;
    const double x = r[0] * it, y = r[1] * it, z = r[2] * it;
    R[0] = c + c1 * x * x; R[1] = c1 * x * y - s * z; R[2] = c1 * x * z + s * y;
    R[3] = c1 * x * y + s * z; R[4] = c + c1 * y * y; R[5] = c1 * y * z - s * x;
    R[6] = c1 * x * z - s * y; R[7] = c1 * y * z + s * x; R[8] = c + c1 * z * z;
}

// evalReprojError (ippe.cpp:748-786): rot2vec -> cv::projectPoints (float output) -> float sums
float reproj_error(const Cam& c, const float* obj /*[4][3]*/, const float* img /*[4][2]*/, const double R[9], const double t[3]) {
    double r[3], Rr[9];
    rot2vec(R, r);
    rodrigues(r, Rr);
    float err = 0;
    for (int i = 0; i < 4; i++) {
        const double X = obj[3 * i], Y = obj[3 * i + 1], Z = obj[3 * i + 2];
        const double x = Rr[0] * X + Rr[1] * Y + Rr[2] * Z + t[0], y = Rr[3] * X + Rr[4] * Y + Rr[5] * Z + t[1];
        double z = Rr[6] * X + Rr[7] * Y + Rr[8] * Z + t[2];
        z = z ? 1. / z : 1;
        const double xn = x * z, yn = y * z;
        const double r2 = xn * xn + yn * yn, r4 = r2 * r2, r6 = r4 * r2;
        const double a1 = 2 * xn * yn, a2 = r2 + 2 * xn * xn, a3 = r2 + 2 * yn * yn;
        const double cdist = 1 + c.k[0] * r2 + c.k[1] * r4 + c.k[4] * r6;
        const double xd = xn * cdist + c.k[2] * a1 + c.k[3] * a2, yd = yn * cdist + c.k[2] * a3 + c.k[3] * a1;
        const float px = (float)(xd * c.fx + c.cx), py = (float)(yd * c.fy + c.cy);
        const float dx = px - img[2 * i], dy = py - img[2 * i + 1];
        err += dx * dx + dy * dy;
    }
    return (float)sqrt(err / (2.0f * 4));
}

}  // namespace

extern "C" {

// pose of one square marker of side `msize` from its four image corners.  cam = fx fy cx cy k1 k2 p1 p2 k3 (the
// reference's float camera matrix and distortion vector, widened).  out = rvec1[3] tvec1[3] err1 rvec2[3] tvec2[3] err2
// with pose 1 the one of smaller reprojection error (doubles; the reference rounds rvec1/tvec1 to float for Marker::Rvec/Tvec
// and keeps err1/err2 as float: both are representable here).
void oracle_ippe_marker_pose(const float* corners, float msize, const double* cam9, double* out14) {
    Cam c;
    c.fx = cam9[0]; c.fy = cam9[1]; c.cx = cam9[2]; c.cy = cam9[3];
    for (int i = 0; i < 5; i++) c.k[i] = cam9[4 + i];
    const float hs = msize / 2.f;                                     // marker.cpp:360
    const float obj[12] = {-hs, hs, 0, hs, hs, 0, hs, -hs, 0, -hs, -hs, 0};
    // undistorted, float-rounded, then widened (ippe.cpp:150, 187-189)
    double ix[4], iy[4];
    for (int i = 0; i < 4; i++) { float x, y; undistort_point(c, corners[2 * i], corners[2 * i + 1], x, y); ix[i] = x; iy[i] = y; }
    // canonical object points: centred (float points stay in the z-plane branch)
    double ox[4], oy[4], xb = 0, yb = 0, zb = 0;
    for (int i = 0; i < 4; i++) { xb += (double)obj[3 * i]; yb += (double)obj[3 * i + 1]; zb += (double)obj[3 * i + 2]; }
    xb /= 4.0; yb /= 4.0; zb /= 4.0;
    for (int i = 0; i < 4; i++) { ox[i] = (double)obj[3 * i] - xb; oy[i] = (double)obj[3 * i + 1] - yb; }
    double H[9];
    homography_ho(ox, oy, ix, iy, H);
    // solveCanonicalForm (ippe.cpp:225-264)
    const double j00 = H[0] - H[6] * H[2], j01 = H[1] - H[7] * H[2], j10 = H[3] - H[6] * H[5], j11 = H[4] - H[7] * H[5];
    double Ra[9], Rb[9], ta[3], tb[3];
    compute_rotations(j00, j01, j10, j11, H[2], H[5], Ra, Rb);
    compute_translation(ox, oy, ix, iy, 4, Ra, ta);
    compute_translation(ox, oy, ix, iy, 4, Rb, tb);
    // Ma = MaCanon * MmodelPoints2Canonical: translation t + R * (-bar)
    double tA[3], tB[3];
    for (int i = 0; i < 3; i++) {
        tA[i] = Ra[3 * i] * (-xb) + Ra[3 * i + 1] * (-yb) + Ra[3 * i + 2] * (-zb) + ta[i];
        tB[i] = Rb[3 * i] * (-xb) + Rb[3 * i + 1] * (-yb) + Rb[3 * i + 2] * (-zb) + tb[i];
    }
    const float ea = reproj_error(c, obj, corners, Ra, tA), eb = reproj_error(c, obj, corners, Rb, tB);
    const bool a_first = ea < eb;                                     // ippe.cpp:793
    const double* R1 = a_first ? Ra : Rb; const double* R2 = a_first ? Rb : Ra;
    const double* t1 = a_first ? tA : tB; const double* t2 = a_first ? tB : tA;
    rot2vec(R1, out14); for (int i = 0; i < 3; i++) out14[3 + i] = t1[i];
    out14[6] = a_first ? ea : eb;
    rot2vec(R2, out14 + 7); for (int i = 0; i < 3; i++) out14[10 + i] = t2[i];
    out14[13] = a_first ? eb : ea;
}

}  // extern "C"
