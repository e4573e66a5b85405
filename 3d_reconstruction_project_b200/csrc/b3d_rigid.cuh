// b3d_rigid.cuh -- small rigid-motion helpers shared by the ICP engine and the global registration: 4x4 product / identity,
// TransformVector6dToMatrix4d, 6x6 LDL^T solve, symmetric 3x3 Jacobi eigen-decomposition, Eigen::umeyama (no scaling).
#pragma once

namespace b3d {

__device__ inline void mat4_mul(const double* A, const double* B, double* C) {
    double R[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += A[4 * i + k] * B[4 * k + j];
            R[4 * i + j] = s;
        }
    for (int i = 0; i < 16; ++i) C[i] = R[i];
}
// TransformVector6dToMatrix4d: R = Rz(x2) Ry(x1) Rx(x0), t = x[3..5]
__device__ inline void vec6_to_mat4(const double* x, double* T) {
    const double ca = cos(x[0]), sa = sin(x[0]);
    const double cb = cos(x[1]), sb = sin(x[1]);
    const double cg = cos(x[2]), sg = sin(x[2]);
    T[0] = cg * cb; T[1] = cg * sb * sa - sg * ca; T[2] = cg * sb * ca + sg * sa; T[3] = x[3];
    T[4] = sg * cb; T[5] = sg * sb * sa + cg * ca; T[6] = sg * sb * ca - cg * sa; T[7] = x[4];
    T[8] = -sb;     T[9] = cb * sa;                T[10] = cb * ca;               T[11] = x[5];
    T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}
// 6x6 symmetric solve by LDL^T without pivoting; false if a pivot is not positive / finite (identity update)
__device__ inline bool solve6(const double* A, const double* b, double* x) {
    double L[36], D[6];
    for (int i = 0; i < 36; ++i) L[i] = 0;
    for (int j = 0; j < 6; ++j) {
        double d = A[6 * j + j];
        for (int k = 0; k < j; ++k) d -= L[6 * j + k] * L[6 * j + k] * D[k];
        if (!(d > 0) || !isfinite(d)) return false;
        D[j] = d;
        L[6 * j + j] = 1;
        for (int i = j + 1; i < 6; ++i) {
            double s = A[6 * i + j];
            for (int k = 0; k < j; ++k) s -= L[6 * i + k] * L[6 * j + k] * D[k];
            L[6 * i + j] = s / d;
        }
    }
    double y[6];
    for (int i = 0; i < 6; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= L[6 * i + k] * y[k];
        y[i] = s;
    }
    for (int i = 0; i < 6; ++i) y[i] /= D[i];
    for (int i = 5; i >= 0; --i) {
        double s = y[i];
        for (int k = i + 1; k < 6; ++k) s -= L[6 * k + i] * x[k];
        x[i] = s;
    }
    for (int i = 0; i < 6; ++i)
        if (!isfinite(x[i])) return false;
    return true;
}
__device__ inline void mat4_identity(double* T) {
    for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
}
// symmetric 3x3 Jacobi eigen-decomposition: A = V diag(w) V^T
__device__ inline void jacobi_eig3(const double* A_in, double* w, double* V) {
    double A[9];
    for (int i = 0; i < 9; ++i) { A[i] = A_in[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 64; ++sweep) {
        const double offd = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
        if (offd == 0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                const double apq = A[3 * p + q];
                if (apq == 0) continue;
                const double theta = (A[3 * q + q] - A[3 * p + p]) / (2 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
                const double c = 1 / sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < 3; ++k) {
                    const double akp = A[3 * k + p], akq = A[3 * k + q];
                    A[3 * k + p] = c * akp - s * akq;
                    A[3 * k + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double apk = A[3 * p + k], aqk = A[3 * q + k];
                    A[3 * p + k] = c * apk - s * aqk;
                    A[3 * q + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[3 * k + p], vkq = V[3 * k + q];
                    V[3 * k + p] = c * vkp - s * vkq;
                    V[3 * k + q] = s * vkp + c * vkq;
                }
            }
    }
    w[0] = A[0]; w[1] = A[4]; w[2] = A[8];
}
__device__ inline double det3(const double* M) {
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}
// Eigen::umeyama(src, dst, with_scaling = false): R = U S V^T of Sigma = cov(dst, src), t = mu_d - R mu_s
__device__ inline void umeyama_from_moments(const double* mu_s, const double* mu_d, const double* Sigma, double* T) {
    double StS[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += Sigma[3 * k + i] * Sigma[3 * k + j];
            StS[3 * i + j] = s;
        }
    double w[3], V[9];
    jacobi_eig3(StS, w, V);
    int ord[3] = {0, 1, 2};
    // sort eigenvalues descending (3 elements)
    if (w[ord[0]] < w[ord[1]]) { int t = ord[0]; ord[0] = ord[1]; ord[1] = t; }
    if (w[ord[1]] < w[ord[2]]) { int t = ord[1]; ord[1] = ord[2]; ord[2] = t; }
    if (w[ord[0]] < w[ord[1]]) { int t = ord[0]; ord[0] = ord[1]; ord[1] = t; }
    double Vs[9], U[9];
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) Vs[3 * r + c] = V[3 * r + ord[c]];
    for (int c = 0; c < 2; ++c) {
        double u[3];
        for (int r = 0; r < 3; ++r) u[r] = Sigma[3 * r] * Vs[c] + Sigma[3 * r + 1] * Vs[3 + c] + Sigma[3 * r + 2] * Vs[6 + c];
        const double nrm = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        if (nrm > 0)
            for (int r = 0; r < 3; ++r) U[3 * r + c] = u[r] / nrm;
        else
            for (int r = 0; r < 3; ++r) U[3 * r + c] = (r == c) ? 1.0 : 0.0;
    }
    {
        double u[3];
        for (int r = 0; r < 3; ++r) u[r] = Sigma[3 * r] * Vs[2] + Sigma[3 * r + 1] * Vs[3 + 2] + Sigma[3 * r + 2] * Vs[6 + 2];
        const double cx = U[3 * 1 + 0] * U[3 * 2 + 1] - U[3 * 2 + 0] * U[3 * 1 + 1];
        const double cy = U[3 * 2 + 0] * U[3 * 0 + 1] - U[3 * 0 + 0] * U[3 * 2 + 1];
        const double cz = U[3 * 0 + 0] * U[3 * 1 + 1] - U[3 * 1 + 0] * U[3 * 0 + 1];
        const double sgn = (u[0] * cx + u[1] * cy + u[2] * cz) < 0 ? -1.0 : 1.0;
        U[2] = sgn * cx; U[5] = sgn * cy; U[8] = sgn * cz;
    }
    double S[3] = {1, 1, 1};
    if (det3(U) * det3(Vs) < 0) S[2] = -1;
    double R[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += U[3 * i + k] * S[k] * Vs[3 * j + k];
            R[3 * i + j] = s;
        }
    mat4_identity(T);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) T[4 * i + j] = R[3 * i + j];
        T[4 * i + 3] = mu_d[i] - (R[3 * i] * mu_s[0] + R[3 * i + 1] * mu_s[1] + R[3 * i + 2] * mu_s[2]);
    }
}

}  // namespace b3d
