// Head pose from 10 facial landmarks, batched on the GPU: replaces the OpenCV calls of
// /root/reference/Skps/core/headpose/pose.py:48-77 get_head_pose():
//   cv2.solvePnP(object_pts, image_pts, K, 0)  -> (rvec, tvec)     [SOLVEPNP_ITERATIVE: DLT start + Levenberg-Marquardt]
//   cv2.projectPoints(reprojectsrc, ...)       -> 8 cube corners
//   cv2.Rodrigues + cv2.decomposeProjectionMatrix -> Euler angles in degrees
// One thread per face, float64.  The start value follows OpenCV's (direct linear transform on normalised image points,
// rotation = orthogonal polar factor), the refinement minimises the pixel reprojection error over (rvec, tvec) with
// Levenberg-Marquardt until the step is below 1e-12 - i.e. the same local minimum OpenCV's 20-iteration solver approaches;
// agreement with cv2 is checked in tests/test_headpose_gpu.py (tolerance stated there).
#include <math.h>

#include "../../include/skps_b200.h"
#include "common.h"

namespace skps {

constexpr int HP_PTS = 10;

__device__ void rodrigues(const double* r, double* R) {
    const double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (th < 1e-12) {
        R[0] = 1; R[1] = -r[2]; R[2] = r[1]; R[3] = r[2]; R[4] = 1; R[5] = -r[0]; R[6] = -r[1]; R[7] = r[0]; R[8] = 1;
        return;
    }
    const double kx = r[0] / th, ky = r[1] / th, kz = r[2] / th, c = cos(th), s = sin(th), v = 1 - c;
    R[0] = c + kx * kx * v; R[1] = kx * ky * v - kz * s; R[2] = kx * kz * v + ky * s;
    R[3] = ky * kx * v + kz * s; R[4] = c + ky * ky * v; R[5] = ky * kz * v - kx * s;
    R[6] = kz * kx * v - ky * s; R[7] = kz * ky * v + kx * s; R[8] = c + kz * kz * v;
}

// rotation matrix -> rotation vector (cv2.Rodrigues inverse)
__device__ void rodrigues_inv(const double* R, double* r) {
    const double cx = R[7] - R[5], cy = R[2] - R[6], cz = R[3] - R[1];
    const double s = 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
    double c = 0.5 * (R[0] + R[4] + R[8] - 1.0);
    c = c > 1 ? 1 : (c < -1 ? -1 : c);
    const double th = acos(c);
    if (s < 1e-5) {
        if (c > 0) { r[0] = r[1] = r[2] = 0; return; }
        // theta = pi
        double t0 = sqrt(fmax((R[0] + 1) * 0.5, 0.0)), t1 = sqrt(fmax((R[4] + 1) * 0.5, 0.0)) * (R[1] < 0 ? -1 : 1);
        double t2 = sqrt(fmax((R[8] + 1) * 0.5, 0.0)) * (R[2] < 0 ? -1 : 1);
        if (fabs(t0) < fabs(t1) && fabs(t0) < fabs(t2) && ((R[5] > 0) != (t1 * t2 > 0))) t2 = -t2;
        const double n = th / sqrt(t0 * t0 + t1 * t1 + t2 * t2);
        r[0] = t0 * n; r[1] = t1 * n; r[2] = t2 * n;
        return;
    }
    const double k = th / (2 * s);
    r[0] = cx * k; r[1] = cy * k; r[2] = cz * k;
}

__device__ double det3(const double* M) {
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}
__device__ void inv3T(const double* M, double* O) {        // O = (M^-1)^T
    const double d = 1.0 / det3(M);
    O[0] = (M[4] * M[8] - M[5] * M[7]) * d; O[1] = (M[5] * M[6] - M[3] * M[8]) * d; O[2] = (M[3] * M[7] - M[4] * M[6]) * d;
    O[3] = (M[2] * M[7] - M[1] * M[8]) * d; O[4] = (M[0] * M[8] - M[2] * M[6]) * d; O[5] = (M[1] * M[6] - M[0] * M[7]) * d;
    O[6] = (M[1] * M[5] - M[2] * M[4]) * d; O[7] = (M[2] * M[3] - M[0] * M[5]) * d; O[8] = (M[0] * M[4] - M[1] * M[3]) * d;
}

// smallest eigenvector of a symmetric 12x12 matrix (cyclic Jacobi); A is destroyed
__device__ void smallest_eigvec12(double* A, double* V, double* out) {
    for (int i = 0; i < 144; ++i) V[i] = (i % 13 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0;
        for (int p = 0; p < 12; ++p) for (int q = p + 1; q < 12; ++q) off += A[p * 12 + q] * A[p * 12 + q];
        if (off < 1e-30) break;
        for (int p = 0; p < 12; ++p)
            for (int q = p + 1; q < 12; ++q) {
                const double apq = A[p * 12 + q];
                if (fabs(apq) < 1e-300) continue;
                const double th = (A[q * 12 + q] - A[p * 12 + p]) / (2 * apq);
                const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1));
                const double c = 1 / sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < 12; ++k) {
                    const double akp = A[k * 12 + p], akq = A[k * 12 + q];
                    A[k * 12 + p] = c * akp - s * akq; A[k * 12 + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 12; ++k) {
                    const double apk = A[p * 12 + k], aqk = A[q * 12 + k];
                    A[p * 12 + k] = c * apk - s * aqk; A[q * 12 + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 12; ++k) {
                    const double vkp = V[k * 12 + p], vkq = V[k * 12 + q];
                    V[k * 12 + p] = c * vkp - s * vkq; V[k * 12 + q] = s * vkp + c * vkq;
                }
            }
    }
    int m = 0;
    for (int i = 1; i < 12; ++i) if (A[i * 12 + i] < A[m * 12 + m]) m = i;
    for (int k = 0; k < 12; ++k) out[k] = V[k * 12 + m];
}

__device__ void project(const double* R, const double* t, const float* X, double f, double cx, double cy, double* uv) {
    const double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
    const double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
    const double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
    uv[0] = f * x / z + cx; uv[1] = f * y / z + cy;
}

__device__ double residual(const double* p, const float* obj, const double* img, double f, double cx, double cy, double* e) {
    double R[9];
    rodrigues(p, R);
    double s = 0;
    for (int i = 0; i < HP_PTS; ++i) {
        double uv[2];
        project(R, p + 3, obj + 3 * i, f, cx, cy, uv);
        e[2 * i] = uv[0] - img[2 * i]; e[2 * i + 1] = uv[1] - img[2 * i + 1];
        s += e[2 * i] * e[2 * i] + e[2 * i + 1] * e[2 * i + 1];
    }
    return s;
}

// cv::RQDecomp3x3 on a rotation matrix -> Euler angles in degrees (what cv2.decomposeProjectionMatrix returns for [R|t])
__device__ void euler_rq(const double* Rin, double* eul) {
    double M[9];
    for (int i = 0; i < 9; ++i) M[i] = Rin[i];
    auto mul = [](const double* A, const double* B, double* C) {
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
    };
    double s = M[7], c = M[8], z = 1.0 / sqrt(c * c + s * s + 2.220446049250313e-16);
    c *= z; s *= z;
    double Qx[9] = {1, 0, 0, 0, c, s, 0, -s, c}, R[9];
    mul(M, Qx, R);
    s = -R[6]; c = R[8]; z = 1.0 / sqrt(c * c + s * s + 2.220446049250313e-16);
    c *= z; s *= z;
    double Qy[9] = {c, 0, -s, 0, 1, 0, s, 0, c};
    mul(R, Qy, M);
    s = M[3]; c = M[4]; z = 1.0 / sqrt(c * c + s * s + 2.220446049250313e-16);
    c *= z; s *= z;
    double Qz[9] = {c, s, 0, -s, c, 0, 0, 0, 1};
    mul(M, Qz, R);
    // decomposition ambiguity: diagonal entries of R (except the last) positive; rotate by 180 degrees where needed
    auto T = [](double* Q) { double t; t = Q[1]; Q[1] = Q[3]; Q[3] = t; t = Q[2]; Q[2] = Q[6]; Q[6] = t; t = Q[5]; Q[5] = Q[7]; Q[7] = t; };
    if (R[0] < 0) {
        if (R[4] < 0) {
            Qz[0] = -Qz[0]; Qz[1] = -Qz[1]; Qz[3] = -Qz[3]; Qz[4] = -Qz[4];
        } else {
            T(Qz);
            Qy[0] = -Qy[0]; Qy[2] = -Qy[2]; Qy[6] = -Qy[6]; Qy[8] = -Qy[8];
        }
    } else if (R[4] < 0) {
        T(Qz); T(Qy);
        Qx[4] = -Qx[4]; Qx[5] = -Qx[5]; Qx[7] = -Qx[7]; Qx[8] = -Qx[8];
    }
    const double deg = 180.0 / 3.14159265358979323846;
    auto ac = [](double v) { return acos(v > 1 ? 1.0 : (v < -1 ? -1.0 : v)); };
    eul[0] = ac(Qx[4]) * (Qx[5] >= 0 ? 1 : -1) * deg;
    eul[1] = ac(Qy[0]) * (Qy[6] >= 0 ? 1 : -1) * deg;
    eul[2] = ac(Qz[0]) * (Qz[1] >= 0 ? 1 : -1) * deg;
}

struct HeadPoseK {
    const float* pts;       // [N][10][2] image points (pixels)
    const float* obj;       // [10][3] model points
    const float* cube;      // [8][3] reprojection source
    int N;
    float f, cx, cy;        // camera: fx = fy = f (pose.py:50: [w,0,w//2; 0,w,h//2; 0,0,1])
    double* rvec; double* tvec; double* euler; double* reproj;   // [N][3], [N][3], [N][3], [N][8][2]
    double* scratch;        // [N][2*144]
};

__global__ void head_pose_kernel(const HeadPoseK k) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= k.N) return;
    const double f = k.f, cx = k.cx, cy = k.cy;
    double img[2 * HP_PTS];
    for (int i = 0; i < 2 * HP_PTS; ++i) img[i] = (double)k.pts[(long long)n * 2 * HP_PTS + i];
    // ---- start value: DLT on normalised image points (OpenCV's non-planar branch), rotation = polar factor
    double* A = k.scratch + (long long)n * 288;
    double* V = A + 144;
    for (int i = 0; i < 144; ++i) A[i] = 0;
    for (int i = 0; i < HP_PTS; ++i) {
        const double X = k.obj[3 * i], Y = k.obj[3 * i + 1], Z = k.obj[3 * i + 2];
        const double x = -(img[2 * i] - cx) / f, y = -(img[2 * i + 1] - cy) / f;
        const double r0[12] = {X, Y, Z, 1, 0, 0, 0, 0, x * X, x * Y, x * Z, x};
        const double r1[12] = {0, 0, 0, 0, X, Y, Z, 1, y * X, y * Y, y * Z, y};
        for (int a = 0; a < 12; ++a) for (int b = 0; b < 12; ++b) A[a * 12 + b] += r0[a] * r0[b] + r1[a] * r1[b];
    }
    double P[12];
    smallest_eigvec12(A, V, P);
    double RR[9] = {P[0], P[1], P[2], P[4], P[5], P[6], P[8], P[9], P[10]};
    double tt[3] = {P[3], P[7], P[11]};
    if (det3(RR) < 0) { for (int i = 0; i < 9; ++i) RR[i] = -RR[i]; for (int i = 0; i < 3; ++i) tt[i] = -tt[i]; }
    double sc = 0;
    for (int i = 0; i < 9; ++i) sc += RR[i] * RR[i];
    sc = sqrt(sc);
    double R[9];
    for (int i = 0; i < 9; ++i) R[i] = RR[i] / (sc / sqrt(3.0));
    for (int it = 0; it < 40; ++it) {                      // Newton iteration to the orthogonal polar factor U V^T
        double iT[9];
        inv3T(R, iT);
        double d = 0;
        for (int i = 0; i < 9; ++i) { const double v = 0.5 * (R[i] + iT[i]); d += fabs(v - R[i]); R[i] = v; }
        if (d < 1e-15) break;
    }
    double p[6];
    rodrigues_inv(R, p);
    for (int i = 0; i < 3; ++i) p[3 + i] = tt[i] * (sqrt(3.0) / sc);
    // ---- Levenberg-Marquardt on the pixel reprojection error (numeric Jacobian, float64)
    double e[2 * HP_PTS], e2[2 * HP_PTS], J[2 * HP_PTS][6];
    double cost = residual(p, k.obj, img, f, cx, cy, e);
    double lambda = 1e-3;
    for (int it = 0; it < 100; ++it) {
        for (int j = 0; j < 6; ++j) {
            const double h = 1e-6 * fmax(1.0, fabs(p[j]));
            double pp[6], pm[6], ep[2 * HP_PTS], em[2 * HP_PTS];
            for (int q = 0; q < 6; ++q) { pp[q] = p[q]; pm[q] = p[q]; }
            pp[j] += h; pm[j] -= h;
            residual(pp, k.obj, img, f, cx, cy, ep);
            residual(pm, k.obj, img, f, cx, cy, em);
            for (int i = 0; i < 2 * HP_PTS; ++i) J[i][j] = (ep[i] - em[i]) / (2 * h);
        }
        double JtJ[36], Jte[6];
        for (int a = 0; a < 6; ++a) {
            Jte[a] = 0;
            for (int i = 0; i < 2 * HP_PTS; ++i) Jte[a] += J[i][a] * e[i];
            for (int b = 0; b < 6; ++b) {
                double s = 0;
                for (int i = 0; i < 2 * HP_PTS; ++i) s += J[i][a] * J[i][b];
                JtJ[a * 6 + b] = s;
            }
        }
        bool improved = false;
        double step = 0;
        for (int tries = 0; tries < 12 && !improved; ++tries) {
            double Mx[36], rhs[6];
            for (int i = 0; i < 36; ++i) Mx[i] = JtJ[i];
            for (int a = 0; a < 6; ++a) { Mx[a * 6 + a] += lambda * fmax(JtJ[a * 6 + a], 1e-12); rhs[a] = -Jte[a]; }
            for (int c = 0; c < 6; ++c) {                  // Gaussian elimination with partial pivoting
                int piv = c;
                for (int r = c + 1; r < 6; ++r) if (fabs(Mx[r * 6 + c]) > fabs(Mx[piv * 6 + c])) piv = r;
                if (piv != c) {
                    for (int q = 0; q < 6; ++q) { const double t = Mx[c * 6 + q]; Mx[c * 6 + q] = Mx[piv * 6 + q]; Mx[piv * 6 + q] = t; }
                    const double t = rhs[c]; rhs[c] = rhs[piv]; rhs[piv] = t;
                }
                const double d = Mx[c * 6 + c];
                for (int r = c + 1; r < 6; ++r) {
                    const double m = Mx[r * 6 + c] / d;
                    for (int q = c; q < 6; ++q) Mx[r * 6 + q] -= m * Mx[c * 6 + q];
                    rhs[r] -= m * rhs[c];
                }
            }
            double dx[6];
            for (int r = 5; r >= 0; --r) {
                double s = rhs[r];
                for (int q = r + 1; q < 6; ++q) s -= Mx[r * 6 + q] * dx[q];
                dx[r] = s / Mx[r * 6 + r];
            }
            double pn[6];
            for (int q = 0; q < 6; ++q) pn[q] = p[q] + dx[q];
            const double c2 = residual(pn, k.obj, img, f, cx, cy, e2);
            if (c2 <= cost) {
                step = 0;
                for (int q = 0; q < 6; ++q) { step += dx[q] * dx[q]; p[q] = pn[q]; }
                for (int i = 0; i < 2 * HP_PTS; ++i) e[i] = e2[i];
                cost = c2; lambda = fmax(lambda * 0.1, 1e-15); improved = true;
            } else {
                lambda *= 10;
            }
        }
        if (!improved || step < 1e-24) break;
    }
    // keep the rotation vector in [0, pi] like cv2.Rodrigues(cv2.Rodrigues(r)) would
    rodrigues(p, R);
    rodrigues_inv(R, p);
    for (int i = 0; i < 3; ++i) { k.rvec[(long long)n * 3 + i] = p[i]; k.tvec[(long long)n * 3 + i] = p[3 + i]; }
    euler_rq(R, k.euler + (long long)n * 3);
    for (int i = 0; i < 8; ++i) project(R, p + 3, k.cube + 3 * i, f, cx, cy, k.reproj + ((long long)n * 8 + i) * 2);
}

}  // namespace skps

using namespace skps;

// Batched get_head_pose (pose.py:48-77).  pts [host] (N,10,2) float32 image points in the order of pose.py:60-61
// (landmarks 17,21,22,26,36,39,42,45,31,35 of a 68-point shape); img_w/img_h the frame size (camera matrix of pose.py:50).
// Outputs [host] float64: rvec (N,3), tvec (N,3), euler (N,3) degrees, reproject (N,8,2).
extern "C" SKPS_API int skps_head_pose(const float* pts, int N, int img_w, int img_h, const float* object_pts,
                                       const float* cube_pts, double* rvec, double* tvec, double* euler, double* reproject) {
    SKPS_CHECK(pts && object_pts && cube_pts && rvec && tvec && euler && reproject && N > 0, "head_pose: bad arguments");
    float *d_pts = nullptr, *d_obj = nullptr, *d_cube = nullptr;
    double* d_out = nullptr;
    const size_t outn = (size_t)N * (3 + 3 + 3 + 16 + 288);
    SKPS_CUDA(cudaMalloc(&d_pts, (size_t)N * 20 * 4));
    SKPS_CUDA(cudaMalloc(&d_obj, 30 * 4));
    SKPS_CUDA(cudaMalloc(&d_cube, 24 * 4));
    SKPS_CUDA(cudaMalloc(&d_out, outn * 8));
    SKPS_CUDA(cudaMemcpy(d_pts, pts, (size_t)N * 20 * 4, cudaMemcpyHostToDevice));
    SKPS_CUDA(cudaMemcpy(d_obj, object_pts, 30 * 4, cudaMemcpyHostToDevice));
    SKPS_CUDA(cudaMemcpy(d_cube, cube_pts, 24 * 4, cudaMemcpyHostToDevice));
    HeadPoseK k;
    k.pts = d_pts; k.obj = d_obj; k.cube = d_cube; k.N = N;
    k.f = (float)img_w; k.cx = (float)(img_w / 2); k.cy = (float)(img_h / 2);
    k.rvec = d_out; k.tvec = d_out + (size_t)N * 3; k.euler = d_out + (size_t)N * 6; k.reproj = d_out + (size_t)N * 9;
    k.scratch = d_out + (size_t)N * 25;
    head_pose_kernel<<<(N + 31) / 32, 32>>>(k);
    SKPS_CUDA(cudaGetLastError());
    SKPS_CUDA(cudaDeviceSynchronize());
    SKPS_CUDA(cudaMemcpy(rvec, k.rvec, (size_t)N * 24, cudaMemcpyDeviceToHost));
    SKPS_CUDA(cudaMemcpy(tvec, k.tvec, (size_t)N * 24, cudaMemcpyDeviceToHost));
    SKPS_CUDA(cudaMemcpy(euler, k.euler, (size_t)N * 24, cudaMemcpyDeviceToHost));
    SKPS_CUDA(cudaMemcpy(reproject, k.reproj, (size_t)N * 128, cudaMemcpyDeviceToHost));
    cudaFree(d_pts); cudaFree(d_obj); cudaFree(d_cube); cudaFree(d_out);
    return 0;
}
