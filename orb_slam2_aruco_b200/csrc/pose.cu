// pose.cu -- B200-native marker pose (sm_100a): IPPE for a square marker, one thread per detected marker.
//
// Behavioural contract: aruco::Marker::calculateExtrinsics (reference Thirdparty/aruco/aruco/marker.cpp:322-343) ->
// aruco::solvePnP -> IPPE::PoseSolver::solveGeneric (Thirdparty/aruco/aruco/ippe.cpp:72-169 and the functions it calls:
// makeCanonicalObjectPoints :647, HomographyHO::homographyHO :912, solveCanonicalForm :225, computeRotations :485,
// computeTranslation :395, sortPosesByReprojError :788, rot2vec :365), plus the second solvePnP of src/Frame.cc:155-177
// whose error ratio err1/err2 < 0.7 flags a marker as good.  Both poses and both reprojection errors are produced at
// once, so that one launch serves detect(image, cameraParams, markerSize) and the Frame constructor's quality test.
// All arithmetic is double in the reference's operation order (no FMA contraction), rounded to float where the
// reference does (normalized points, projected points, error sums, outputs); results agree with the CPU oracle
// (oracle/ippe_oracle.cpp) to ~1e-9 (only acos / sin / cos / hypot differ in the last bit between libm and CUDA).
#include "common.h"
#include <math.h>
#include <float.h>

namespace b200 {

struct PoseCam { double fx, fy, cx, cy, k[5]; };

__device__ void pose_eigen_sym3(const double* Ain, double* W, double* V) {      // eigenvalues descending, eigenvectors in rows (cv::eigen)
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
#define B200_ROT(v0, v1) { const double a0 = v0, b0 = v1; v0 = a0 * c - b0 * s; v1 = a0 * s + b0 * c; }
        for (int i = 0; i < k; i++) B200_ROT(A[i][k], A[i][l]);
        for (int i = k + 1; i < l; i++) B200_ROT(A[k][i], A[i][l]);
        for (int i = l + 1; i < 3; i++) B200_ROT(A[k][i], A[l][i]);
        for (int i = 0; i < 3; i++) B200_ROT(V[3 * k + i], V[3 * l + i]);
#undef B200_ROT
    }
    for (int i = 0; i < 3; i++) W[i] = A[i][i];
    for (int k = 0; k < 2; k++) {
        int m = k;
        for (int i = k + 1; i < 3; i++) if (W[m] < W[i]) m = i;
        if (m != k) {
            const double tw = W[m]; W[m] = W[k]; W[k] = tw;
            for (int i = 0; i < 3; i++) { const double tv = V[3 * m + i]; V[3 * m + i] = V[3 * k + i]; V[3 * k + i] = tv; }
        }
    }
}

__device__ void pose_normalize(const double* xs, const double* ys, double* D0, double* D1, double* T, double* Ti) {     // n = 4, ippe.cpp:809-910
    double xm = 0, ym = 0;
    for (int i = 0; i < 4; i++) { xm += xs[i]; ym += ys[i]; }
    xm /= 4.0; ym /= 4.0;
    double kappa = 0;
    for (int i = 0; i < 4; i++) { D0[i] = xs[i] - xm; D1[i] = ys[i] - ym; kappa += D0[i] * D0[i] + D1[i] * D1[i]; }
    const double beta = sqrt(2 * 4 / kappa);
    for (int i = 0; i < 4; i++) { D0[i] *= beta; D1[i] *= beta; }
    for (int i = 0; i < 9; i++) { T[i] = 0; Ti[i] = 0; }
    T[0] = 1.0 / beta; T[4] = 1.0 / beta; T[2] = xm; T[5] = ym; T[8] = 1;
    Ti[0] = beta; Ti[4] = beta; Ti[2] = -beta * xm; Ti[5] = -beta * ym; Ti[8] = 1;
}

__device__ void pose_mul3(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += A[3 * i + k] * B[3 * k + j]; C[3 * i + j] = s; }
}

__device__ void pose_homography(const double* ax, const double* ay, const double* bx, const double* by, double* H) {     // ippe.cpp:912-1033
    const int n = 4;
    double A0[4], A1[4], B0[4], B1[4], TA[9], TAi[9], TB[9], TBi[9];
    pose_normalize(ax, ay, A0, A1, TA, TAi);
    pose_normalize(bx, by, B0, B1, TB, TBi);
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
    double g00 = 0, g01 = 0, g11 = 0;
    for (int i = 0; i < n; i++) { g00 += A0[i] * A0[i]; g01 += A0[i] * A1[i]; g11 += A1[i] * A1[i]; }
    const double dt = g00 * g11 - g01 * g01;
    const double i00 = g11 / dt, i01 = -g01 / dt, i10 = -g01 / dt, i11 = g00 / dt;
    double Pp0[4], Pp1[4];
    for (int i = 0; i < n; i++) { Pp0[i] = i00 * A0[i] + i01 * A1[i]; Pp1[i] = i10 * A0[i] + i11 * A1[i]; }
    double Bx[2][3], By[2][3];
    for (int j = 0; j < 3; j++) {
        double s0 = 0, s1 = 0, t0 = 0, t1 = 0;
        for (int i = 0; i < n; i++) { s0 += Pp0[i] * Mx[i][j]; s1 += Pp1[i] * Mx[i][j]; t0 += Pp0[i] * My[i][j]; t1 += Pp1[i] * My[i][j]; }
        Bx[0][j] = s0; Bx[1][j] = s1; By[0][j] = t0; By[1][j] = t1;
    }
    double DDT[9];
    for (int i = 0; i < 9; i++) DDT[i] = 0;
    // D rows are consumed as they are produced, in row order 0..7 (the order the reference's D^T * D accumulates them)
    for (int half = 0; half < 2; half++)
        for (int i = 0; i < n; i++) {
            double d[3];
            for (int j = 0; j < 3; j++) {
                const double e = half == 0 ? A0[i] * Bx[0][j] + A1[i] * Bx[1][j] : A0[i] * By[0][j] + A1[i] * By[1][j];
                d[j] = (half == 0 ? Mx[i][j] : My[i][j]) - e;
            }
            for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) DDT[3 * a + b] += d[a] * d[b];
        }
    double W[3], U[9];
    pose_eigen_sym3(DDT, W, U);
    const double h7 = U[6], h8 = U[7], h9 = U[8];
    const double h1 = -(Bx[0][0] * h7 + Bx[0][1] * h8 + Bx[0][2] * h9), h2 = -(Bx[1][0] * h7 + Bx[1][1] * h8 + Bx[1][2] * h9);
    const double h4 = -(By[0][0] * h7 + By[0][1] * h8 + By[0][2] * h9), h5 = -(By[1][0] * h7 + By[1][1] * h8 + By[1][2] * h9);
    const double h3 = -(mC1 * h7 + mC2 * h8), h6 = -(mC3 * h7 + mC4 * h8);
    const double Hn[9] = {h1, h2, h3, h4, h5, h6, h7, h8, h9};
    double T1[9];
    pose_mul3(TB, Hn, T1);
    pose_mul3(T1, TAi, H);
    const double s = H[8];
    for (int i = 0; i < 9; i++) H[i] = H[i] / s;
}

__device__ void pose_rotations(double j00, double j01, double j10, double j11, double p, double q, double* R1, double* R2) {     // ippe.cpp:485-590, 1036-1080
    double Ra[9];
    {
        double ax = p, ay = q, az = 1.0;
        const double nrm = sqrt(ax * ax + ay * ay + az * az);
        ax /= nrm; ay /= nrm; az /= nrm;
        const double c = az;
        if (fabs(1.0 + c) < (double)FLT_EPSILON) { for (int i = 0; i < 9; i++) Ra[i] = 0; Ra[0] = 1; Ra[4] = 1; Ra[8] = -1; }
        else {
            const double d = 1.0 / (1.0 + c), ax2 = ax * ax, ay2 = ay * ay, axay = ax * ay;
            Ra[0] = -ax2 * d + 1.0; Ra[1] = -axay * d; Ra[2] = -ax;
            Ra[3] = -axay * d; Ra[4] = -ay2 * d + 1.0; Ra[5] = -ay;
            Ra[6] = ax; Ra[7] = ay; Ra[8] = 1.0 - (ax2 + ay2) * d;
        }
    }
    const double rv00 = Ra[0], rv01 = Ra[3], rv02 = Ra[6], rv10 = Ra[1], rv11 = Ra[4], rv12 = Ra[7], rv20 = Ra[2], rv21 = Ra[5], rv22 = Ra[8];   // Rv = Ra^T
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

__device__ void pose_translation(const double* ox, const double* oy, const double* ix, const double* iy, const double* R, double* t) {      // ippe.cpp:395-483
    const double ATA00 = 4, ATA11 = 4;
    double ATA02 = 0, ATA12 = 0, ATA20 = 0, ATA21 = 0, ATA22 = 0, ATb0 = 0, ATb1 = 0, ATb2 = 0;
    for (int i = 0; i < 4; i++) {
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

__device__ void pose_rot2vec(const double* R, double* r) {      // ippe.cpp:365-393
    const double trace = R[0] + R[4] + R[8];
    const double w_norm = acos((trace - 1.0) / 2.0);
    const double d = 1 / (2 * sin(w_norm)) * w_norm;
    if (w_norm < (double)FLT_EPSILON) { r[0] = r[1] = r[2] = 0; return; }
    r[0] = d * (R[7] - R[5]); r[1] = d * (R[2] - R[6]); r[2] = d * (R[3] - R[1]);
}

// evalReprojError (ippe.cpp:748-786): rot2vec -> cv::Rodrigues -> cv::projectPoints (float output) -> float sums
__device__ float pose_reproj_error(const PoseCam& c, const float* obj, const float* img, const double* R, const double* t) {
    double r[3], Rr[9];
    pose_rot2vec(R, r);
    const double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (theta < DBL_EPSILON) { for (int i = 0; i < 9; i++) Rr[i] = 0; Rr[0] = Rr[4] = Rr[8] = 1; }
    else {
        const double cs = cos(theta), sn = sin(theta), c1 = 1. - cs, it = 1. / theta;
        const double x = r[0] * it, y = r[1] * it, z = r[2] * it;
        Rr[0] = cs + c1 * x * x; Rr[1] = c1 * x * y - sn * z; Rr[2] = c1 * x * z + sn * y;
        Rr[3] = c1 * x * y + sn * z; Rr[4] = cs + c1 * y * y; Rr[5] = c1 * y * z - sn * x;
        Rr[6] = c1 * x * z - sn * y; Rr[7] = c1 * y * z + sn * x; Rr[8] = cs + c1 * z * z;
    }
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
        const float dx = __fsub_rn(px, img[2 * i]), dy = __fsub_rn(py, img[2 * i + 1]);
        err = __fadd_rn(err, __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    }
    return (float)sqrt((double)__fdiv_rn(err, 8.0f));
}

__global__ void __launch_bounds__(64)
k_pose(const b200_marker* __restrict__ markers, const int* __restrict__ counts, int n_batch, int cap, float msize, PoseCam c,
       b200_marker_pose* __restrict__ poses) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_batch * cap) return;
    const int f = i / cap, m = i - f * cap;
    if (m >= min(counts[f], cap)) return;
    const b200_marker mk = markers[i];
    const float hs = __fdiv_rn(msize, 2.f);                                     // marker.cpp:360
    const float obj[12] = {-hs, hs, 0, hs, hs, 0, hs, -hs, 0, -hs, -hs, 0};
    // cv::undistortPoints: 5 fixed-point iterations, stored as float, widened again (ippe.cpp:150, 187-189)
    double ix[4], iy[4];
    const bool has_dist = c.k[0] != 0 || c.k[1] != 0 || c.k[2] != 0 || c.k[3] != 0 || c.k[4] != 0;
    for (int p = 0; p < 4; p++) {
        double x = ((double)mk.xy[2 * p] - c.cx) * (1.0 / c.fx), y = ((double)mk.xy[2 * p + 1] - c.cy) * (1.0 / c.fy);
        const double x0 = x, y0 = y;
        if (has_dist)
            for (int j = 0; j < 5; j++) {
                const double r2 = x * x + y * y;
                const double icdist = 1.0 / (1 + ((c.k[4] * r2 + c.k[1]) * r2 + c.k[0]) * r2);
                if (icdist < 0) { x = x0; y = y0; break; }
                const double dx = 2 * c.k[2] * x * y + c.k[3] * (r2 + 2 * x * x);
                const double dy = c.k[2] * (r2 + 2 * y * y) + 2 * c.k[3] * x * y;
                x = (x0 - dx) * icdist;
                y = (y0 - dy) * icdist;
            }
        ix[p] = (double)(float)x; iy[p] = (double)(float)y;
    }
    // canonical object points: centred (with float object points the reference never leaves the z-plane branch, ippe.cpp:675-679)
    double ox[4], oy[4], xb = 0, yb = 0, zb = 0;
    for (int p = 0; p < 4; p++) { xb += (double)obj[3 * p]; yb += (double)obj[3 * p + 1]; zb += (double)obj[3 * p + 2]; }
    xb /= 4.0; yb /= 4.0; zb /= 4.0;
    for (int p = 0; p < 4; p++) { ox[p] = (double)obj[3 * p] - xb; oy[p] = (double)obj[3 * p + 1] - yb; }
    double H[9];
    pose_homography(ox, oy, ix, iy, H);
    const double j00 = H[0] - H[6] * H[2], j01 = H[1] - H[7] * H[2], j10 = H[3] - H[6] * H[5], j11 = H[4] - H[7] * H[5];
    double Ra[9], Rb[9], ta[3], tb[3];
    pose_rotations(j00, j01, j10, j11, H[2], H[5], Ra, Rb);
    pose_translation(ox, oy, ix, iy, Ra, ta);
    pose_translation(ox, oy, ix, iy, Rb, tb);
    double tA[3], tB[3];
    for (int q = 0; q < 3; q++) {
        tA[q] = Ra[3 * q] * (-xb) + Ra[3 * q + 1] * (-yb) + Ra[3 * q + 2] * (-zb) + ta[q];
        tB[q] = Rb[3 * q] * (-xb) + Rb[3 * q + 1] * (-yb) + Rb[3 * q + 2] * (-zb) + tb[q];
    }
    const float ea = pose_reproj_error(c, obj, mk.xy, Ra, tA), eb = pose_reproj_error(c, obj, mk.xy, Rb, tB);
    const bool a_first = ea < eb;                                               // ippe.cpp:793
    double r1[3], r2[3];
    pose_rot2vec(a_first ? Ra : Rb, r1);
    pose_rot2vec(a_first ? Rb : Ra, r2);
    b200_marker_pose o;
    for (int q = 0; q < 3; q++) {
        o.rvec[q] = (float)r1[q]; o.tvec[q] = (float)(a_first ? tA[q] : tB[q]);
        o.rvec2[q] = (float)r2[q]; o.tvec2[q] = (float)(a_first ? tB[q] : tA[q]);
    }
    o.err1 = a_first ? ea : eb; o.err2 = a_first ? eb : ea;
    poses[i] = o;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_aruco_pose(const b200_marker* markers, const int32_t* counts, int n_batch, int marker_cap, float marker_size,
                    const float* cam9, b200_marker_pose* poses, int device, void* stream) {
    if (n_batch < 0 || marker_cap < 0) return fail(B200_EINVAL, "negative %s", "size");
    if (!(marker_size > 0)) return fail(B200_EINVAL, "markerSize<=0: invalid %s", "markerSize");             // marker.cpp:328
    if (!cam9 || !(cam9[0] != 0) || !(cam9[1] != 0)) return fail(B200_EINVAL, "invalid camera %s", "parameters");     // marker.cpp:309
    DeviceScope _ds; int rc = use_device(device);
    if (rc) return rc;
    if (n_batch == 0 || marker_cap == 0) return B200_OK;
    if (!markers || !counts || !poses) return fail(B200_EINVAL, "null %s", "pointer");
    PoseCam c;
    c.fx = cam9[0]; c.fy = cam9[1]; c.cx = cam9[2]; c.cy = cam9[3];
    for (int i = 0; i < 5; i++) c.k[i] = cam9[4 + i];
    const int total = n_batch * marker_cap;
    B200_LAUNCH(k_pose, (total + 63) / 64, 64, 0, (cudaStream_t)stream, markers, counts, n_batch, marker_cap, marker_size, c, poses);
    B200_CUDA(cudaGetLastError());
    return B200_OK;
}

int b200_aruco_pose_host(const b200_marker* markers, int n_markers, float marker_size, const float* cam9, b200_marker_pose* poses, int device) {
    if (n_markers < 0) return fail(B200_EINVAL, "negative %s", "size");
    DeviceScope _ds; int rc = use_device(device);
    if (rc) return rc;
    if (n_markers == 0) return B200_OK;
    if (!markers || !poses) return fail(B200_EINVAL, "null %s", "pointer");
    cudaStream_t ts = nullptr;
    if ((rc = host_call_stream(device, &ts))) return rc;
    DevBuf dm, dc, dp;
    if ((rc = dm.upload(markers, sizeof(b200_marker) * n_markers)) || (rc = dc.upload(&n_markers, 4)) || (rc = dp.alloc(sizeof(b200_marker_pose) * n_markers))) return rc;
    if ((rc = b200_aruco_pose((const b200_marker*)dm.p, (const int32_t*)dc.p, 1, n_markers, marker_size, cam9, (b200_marker_pose*)dp.p, device, ts))) return rc;
    B200_D2H(poses, dp.p, sizeof(b200_marker_pose) * n_markers);
    return B200_OK;
}

}  // extern "C"
