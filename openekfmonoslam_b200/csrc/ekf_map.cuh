// ekf_map.cuh -- device-side map management: the state vector and the covariance never leave HBM.
//
//   k_map_plan     one CTA per filter: which features go (removeBadMapFeatures, E/MapManagement.cpp:279-307; the unseen
//                  ones under the policy of E/EKF.cpp:575-589), which single inverse-depth feature becomes XYZ
//                  (convertMapFeaturesInverseDepthToDepth, E/MapManagement.cpp:311-341,501-524), the new layout
//                  (ordered scan = the reference's covarianceMatrixPos bookkeeping, :238-256,:486-496), the compacted
//                  per-feature arrays and state vector, and the new-row -> old-row table
//   k_map_gather   P2[i][j] = P[src(i)][src(j)]: removeRowsAndColumnsFromMat (E/MapManagement.cpp:117-208) as one
//                  gather pass; HBM-bound, 16 n^2 bytes
//   k_map_convert  the three rows / columns of the converted feature: J P[6 rows] and J P66 J^T (E/MapManagement.cpp:345-455)
//   k_map_commit   the new sizes become the live ones (the host swaps the buffer pointers)
//   k_add_prepare / k_add_cov   addFeaturesToStateAndCovariance (E/AddMapFeature.cpp:116-366) for a batch of new
//                  inverse-depth features in two launches: all new rows are J_a P[0:7, :], their mirrors, and the
//                  new-vs-new blocks J_a P77 J_b^T (+ the measurement-noise block on the diagonal)
// P stays exactly symmetric: every kernel writes an entry and its mirror from the same value.
#pragma once

#include "ekf_kernels.cuh"

namespace ekf {

struct MapPolicy {   // ekfb_map_policy
    int min_matches, max_features, max_size, always_remove_unseen;
    double good_pct, linearity_thr;
};

// computeLinearityIndex (E/MapManagement.cpp:311-341)
__device__ inline double linearity_index(const double* r, const double* y, double Prr)
{
    const double sigma = sqrt(Prr) / (y[5] * y[5]);
    double m[3], toCam[3], toFirst[3];
    direction(y[3], y[4], m);
    double dot = 0.0, d1 = 0.0, d2 = 0.0;
    for (int i = 0; i < 3; ++i) {
        const double p = y[i] + m[i] / y[5];
        toCam[i] = p - r[i];
        toFirst[i] = p - y[i];
        dot += toCam[i] * toFirst[i];
        d1 += toFirst[i] * toFirst[i];
        d2 += toCam[i] * toCam[i];
    }
    d1 = sqrt(d1);
    d2 = sqrt(d2);
    return 4.0 * sigma * (dot / (d1 * d2)) / d2;
}

__global__ void __launch_bounds__(256) k_map_plan(DevView v, MapPolicy pol)
{
    const int f = blockIdx.x, tid = threadIdx.x;
    int* dm = fdims(v, f);
    const int N = dm[D_N_FEAT], n = dm[D_N_STATE];
    const size_t fo = (size_t)f * v.Nmax;
    const double* x = v.x + (size_t)f * v.ld;
    double* x2 = v.x2 + (size_t)f * v.ld;
    const double* P = v.P + (size_t)f * v.nmax * v.ld;
    int* rowsrc = v.rowsrc + (size_t)f * v.nmax;
    uint8_t* flag = v.mapflag + fo;
    __shared__ int sCnt[2], sConv, sScanN[256], sScanD[256], sCarryN, sCarryD;
    if (tid == 0) { sCnt[0] = sCnt[1] = 0; sConv = 0x7fffffff; sCarryN = 0; sCarryD = 13; }
    __syncthreads();
    // (1) removeBadMapFeatures: float ratio; 0/0 (never predicted) and x/0 compare false
    int nb = 0, db = 0;
    for (int i = tid; i < N; i += blockDim.x) {
        const float pct = (float)(unsigned)v.tmatch[fo + i] / (float)(unsigned)v.tpred[fo + i];
        const bool bad = (double)pct < pol.good_pct;
        flag[i] = bad ? 1 : 0;
        if (bad) { nb++; db += v.ftype[fo + i] == kTypeInvDepth ? 6 : 3; }
    }
    if (nb) { atomicAdd(&sCnt[0], nb); atomicAdd(&sCnt[1], db); }
    __syncthreads();
    const int nBad = sCnt[0], rowsAfterBad = n - sCnt[1];
    const int needed = pol.min_matches - (dm[D_N_INL] + dm[D_N_RESC]);
    // (2) E/EKF.cpp:580-589: unsigned compare of map size + needed (needed > 0 here) against the limits
    const bool dropUnseen = needed > 0 && (pol.always_remove_unseen ||
                                           (pol.max_features > 0 && (N - nBad) + needed > pol.max_features) ||
                                           (pol.max_size > 0 && rowsAfterBad + needed * 6 > pol.max_size));
    __syncthreads();
    if (tid == 0) sCnt[0] = 0;
    __syncthreads();
    if (dropUnseen) {
        int nu = 0;
        for (int i = tid; i < N; i += blockDim.x)
            if (!flag[i] && !v.vis[fo + i]) { flag[i] = 2; nu++; }
        if (nu) atomicAdd(&sCnt[0], nu);
    }
    __syncthreads();
    const int nUnseen = sCnt[0];
    // (3) first surviving inverse-depth feature whose linearity index is under the threshold
    for (int i = tid; i < N; i += blockDim.x) {
        if (flag[i] || v.ftype[fo + i] != kTypeInvDepth) continue;
        const int o = v.foff[fo + i];
        if (linearity_index(x, x + o, P[(size_t)(o + 5) * v.ld + o + 5]) < pol.linearity_thr) atomicMin(&sConv, i);
    }
    __syncthreads();
    const int conv = sConv == 0x7fffffff ? -1 : sConv;
    // (4) ordered scan -> new feature index and covarianceMatrixPos; compacted copies
    if (tid < 13) { rowsrc[tid] = tid; x2[tid] = x[tid]; }
    for (int base = 0; base < N; base += blockDim.x) {
        const int i = base + tid;
        int keep = 0, d = 0, dOld = 0, o = 0;
        if (i < N) {
            keep = flag[i] == 0;
            dOld = v.ftype[fo + i] == kTypeInvDepth ? 6 : 3;
            d = keep ? (i == conv ? 3 : dOld) : 0;
            o = v.foff[fo + i];
        }
        sScanN[tid] = keep;
        sScanD[tid] = d;
        __syncthreads();
        for (int s = 1; s < 256; s <<= 1) {   // Hillis-Steele inclusive scan
            const int a = tid >= s ? sScanN[tid - s] : 0, b = tid >= s ? sScanD[tid - s] : 0;
            __syncthreads();
            sScanN[tid] += a;
            sScanD[tid] += b;
            __syncthreads();
        }
        const int ni = sCarryN + sScanN[tid] - keep, no = sCarryD + sScanD[tid] - d;
        if (keep) {
            v.foff2[fo + ni] = no;
            v.tpred2[fo + ni] = v.tpred[fo + i];
            v.tmatch2[fo + ni] = v.tmatch[fo + i];
            const uint32_t* ds = reinterpret_cast<const uint32_t*>(v.desc + (fo + i) * 32);
            uint32_t* dd = reinterpret_cast<uint32_t*>(v.desc2 + (fo + ni) * 32);
            for (int a = 0; a < 8; ++a) dd[a] = ds[a];
            for (int a = 0; a < d; ++a) rowsrc[no + a] = o + a;
            if (i == conv) {   // convertToDepth: XYZ point and the 3x6 Jacobian (E/MapManagement.cpp:347-390)
                const double* y = x + o;
                const double th = y[3], ph = y[4], rho = y[5];
                double m[3];
                direction(th, ph, m);
                for (int a = 0; a < 3; ++a) x2[no + a] = y[a] + m[a] / rho;
                double* J = v.convJ + (size_t)f * 18;
                for (int a = 0; a < 18; ++a) J[a] = 0.0;
                J[0] = J[7] = J[14] = 1.0;
                J[3] = cos(ph) * cos(th) / rho;   J[15] = -cos(ph) * sin(th) / rho;
                J[4] = -sin(ph) * sin(th) / rho;  J[10] = -cos(ph) / rho;  J[16] = -sin(ph) * cos(th) / rho;
                J[5] = -m[0] / (rho * rho);       J[11] = -m[1] / (rho * rho);  J[17] = -m[2] / (rho * rho);
                v.ftype2[fo + ni] = kTypeXYZ;
                dm[D_MAP_CONVERT] = ni;
                dm[D_MAP_CONV_OLDOFF] = o;
                dm[D_MAP_CONV_NEWOFF] = no;
            } else {
                v.ftype2[fo + ni] = v.ftype[fo + i];
                for (int a = 0; a < d; ++a) x2[no + a] = x[o + a];
            }
        }
        __syncthreads();
        if (tid == 255) { sCarryN += sScanN[255]; sCarryD += sScanD[255]; }
        __syncthreads();
    }
    if (tid == 0) {
        if (conv < 0) dm[D_MAP_CONVERT] = -1;
        dm[D_MAP_NEW_N] = sCarryD;
        dm[D_MAP_NEW_NF] = sCarryN;
        dm[D_MAP_NBAD] = nBad;
        dm[D_MAP_NUNSEEN] = nUnseen;
        dm[D_MAP_NEEDED] = needed;
        dm[D_MAP_CHANGED] = (nBad + nUnseen > 0 || conv >= 0) ? 1 : 0;
    }
}

// grid (ceil(nmaxNew / 32), ceil(nmaxNew / 32), F), block (32, 8)
__global__ void __launch_bounds__(256) k_map_gather(DevView v)
{
    const int f = blockIdx.z;
    const int* dm = fdims(v, f);
    const int nn = dm[D_MAP_NEW_N];
    const int j = blockIdx.x * 32 + threadIdx.x;
    if (blockIdx.x * 32 >= nn || blockIdx.y * 32 >= nn || j >= nn) return;
    const int* rowsrc = v.rowsrc + (size_t)f * v.nmax;
    const double* P = v.P + (size_t)f * v.nmax * v.ld;
    double* P2 = v.P2 + (size_t)f * v.nmax * v.ld;
    const int sj = rowsrc[j];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = blockIdx.y * 32 + threadIdx.y + 8 * r;
        if (i < nn) P2[(size_t)i * v.ld + j] = P[(size_t)rowsrc[i] * v.ld + sj];
    }
}

// grid (ceil(nmax / 256), F)
__global__ void __launch_bounds__(256) k_map_convert(DevView v)
{
    const int f = blockIdx.y;
    const int* dm = fdims(v, f);
    if (dm[D_MAP_CONVERT] < 0) return;
    const int nn = dm[D_MAP_NEW_N], c0 = dm[D_MAP_CONV_NEWOFF], o0 = dm[D_MAP_CONV_OLDOFF];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nn) return;
    const int* rowsrc = v.rowsrc + (size_t)f * v.nmax;
    const double* P = v.P + (size_t)f * v.nmax * v.ld;
    double* P2 = v.P2 + (size_t)f * v.nmax * v.ld;
    const double* J = v.convJ + (size_t)f * 18;
    if (j >= c0 && j < c0 + 3) {
        if (j != c0) return;
        double T[18], M[9];   // T = J P66, M = T J^T
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 6; ++b) {
                double s = 0.0;
                for (int k = 0; k < 6; ++k) s += J[a * 6 + k] * P[(size_t)(o0 + k) * v.ld + o0 + b];
                T[a * 6 + b] = s;
            }
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                double s = 0.0;
                for (int k = 0; k < 6; ++k) s += T[a * 6 + k] * J[b * 6 + k];
                M[a * 3 + b] = s;
            }
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b <= a; ++b) {
                P2[(size_t)(c0 + a) * v.ld + c0 + b] = M[a * 3 + b];
                P2[(size_t)(c0 + b) * v.ld + c0 + a] = M[a * 3 + b];
            }
        return;
    }
    const int sj = rowsrc[j];
    double col[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) col[k] = P[(size_t)(o0 + k) * v.ld + sj];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) s += J[a * 6 + k] * col[k];
        P2[(size_t)(c0 + a) * v.ld + j] = s;
        P2[(size_t)j * v.ld + c0 + a] = s;
    }
}

__global__ void k_map_commit(DevView v)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= v.F) return;
    int* dm = fdims(v, f);
    dm[D_N_STATE] = dm[D_MAP_NEW_N];
    dm[D_N_FEAT] = dm[D_MAP_NEW_NF];
}

// ---------------------------------------------------------------------------------------------
// addFeaturesToStateAndCovariance for `cnt` new inverse-depth features of filter f observed at pixels uv.
// k_add_prepare: one warp per new feature a: the feature (r, theta, phi, rho0) (E/AddMapFeature.cpp:293-334), the
// 6x7 Jacobian with respect to (r, q) and the 6x6 noise block JH diag(sx^2, sy^2, srho^2) JH^T (:116-216,:236-239),
// descriptor and counters.  addJ[a] = 42 + 36 doubles.
// ---------------------------------------------------------------------------------------------
constexpr int kAddJ = 78;

__global__ void __launch_bounds__(32) k_add_prepare(DevView v, int f, int n0, int N0, int cnt, double px_err_x, double px_err_y,
                                                    double rho0, double rho_sd)
{
    const int a = blockIdx.x, lane = threadIdx.x;
    const size_t fo = (size_t)f * v.Nmax;
    double* x = v.x + (size_t)f * v.ld;
    const CamParams& c = v.cam;
    if (lane < 8)
        reinterpret_cast<uint32_t*>(v.desc + (fo + N0 + a) * 32)[lane] = reinterpret_cast<const uint32_t*>(v.adddesc + (size_t)a * 32)[lane];
    if (lane != 0) return;
    const double* uv = v.adduv + 2 * a;
    const double pxm = uv[0] - c.cx, pym = uv[1] - c.cy;
    const double mx = c.dx * pxm, my = c.dy * pym, rd = mx * mx + my * my;
    const double dist = 1 + c.k1 * rd + c.k2 * rd * rd;
    const double und[2] = {c.cx + pxm * dist, c.cy + pym * dist};
    const double* q = x + 3;
    double R[9];
    quat_to_rot(q, R);
    const double gc[3] = {-(c.cx - und[0]) / c.fx, -(c.cy - und[1]) / c.fy, 1.0};
    double gw[3];
    mat3_vec(R, gc, gw);
    const double xw = gw[0], yw = gw[1], zw = gw[2];
    double* y = x + n0 + 6 * a;
    y[0] = x[0]; y[1] = x[1]; y[2] = x[2];
    y[3] = atan2(xw, zw);
    y[4] = atan2(-yw, sqrt(xw * xw + zw * zw));
    y[5] = rho0;
    v.ftype[fo + N0 + a] = kTypeInvDepth;
    v.foff[fo + N0 + a] = n0 + 6 * a;
    v.tpred[fo + N0 + a] = 0;
    v.tmatch[fo + N0 + a] = 0;
    const double xxzz = xw * xw + zw * zw, sq = sqrt(xxzz), nsq = xxzz + yw * yw;
    const double dth[3] = {zw / xxzz, 0.0, -xw / xxzz};
    const double dph[3] = {xw * yw / (nsq * sq), -sq / nsq, zw * yw / (nsq * sq)};
    double dgw_dq[12];
    drot_dq(q, gc, dgw_dq);
    double* J = v.addJ + (size_t)a * kAddJ;
    for (int i = 0; i < kAddJ; ++i) J[i] = 0.0;
    J[0] = J[8] = J[16] = 1.0;
    for (int i = 0; i < 4; ++i) {
        double s = 0.0, t = 0.0;
        for (int k = 0; k < 3; ++k) { s += dth[k] * dgw_dq[k * 4 + i]; t += dph[k] * dgw_dq[k * 4 + i]; }
        J[3 * 7 + 3 + i] = s;
        J[4 * 7 + 3 + i] = t;
    }
    double sub[6];
    for (int j = 0; j < 3; ++j) {
        double s = 0.0, t = 0.0;
        for (int k = 0; k < 3; ++k) { s += dth[k] * R[k * 3 + j]; t += dph[k] * R[k * 3 + j]; }
        sub[j] = s; sub[3 + j] = t;
    }
    const double s2[4] = {sub[0] * (1.0 / c.fx), sub[1] * (1.0 / c.fy), sub[3] * (1.0 / c.fx), sub[4] * (1.0 / c.fy)};
    const double k12 = c.k1 + 2.0 * c.k2 * rd, k1p = 1.0 + c.k1 * rd + c.k2 * rd * rd;
    const double dx2 = 2.0 * c.dx * c.dx, dy2 = 2.0 * c.dy * c.dy;
    const double dhu[4] = {k1p + pxm * k12 * (pxm * dx2), pxm * k12 * (pym * dy2), pym * k12 * (pxm * dx2),
                           pym * k12 * (pym * dy2) + k1p};
    double JH[18] = {0};
    JH[9] = s2[0] * dhu[0] + s2[1] * dhu[2];  JH[10] = s2[0] * dhu[1] + s2[1] * dhu[3];
    JH[12] = s2[2] * dhu[0] + s2[3] * dhu[2]; JH[13] = s2[2] * dhu[1] + s2[3] * dhu[3];
    JH[17] = 1.0;
    const double noise[3] = {px_err_x * px_err_x, px_err_y * px_err_y, rho_sd * rho_sd};
    for (int r = 0; r < 6; ++r)
        for (int s = 0; s < 6; ++s) {
            double b = 0.0;
            for (int k = 0; k < 3; ++k) b += JH[r * 3 + k] * noise[k] * JH[s * 3 + k];
            J[42 + r * 6 + s] = b;
        }
    if (a == 0) {
        int* dm = fdims(v, f);
        dm[D_N_STATE] = n0 + 6 * cnt;
        dm[D_N_FEAT] = N0 + cnt;
    }
}

// grid (ceil((n0 + 6 cnt) / 256), cnt), block 256: block row a owns the six new rows of feature a.
__global__ void __launch_bounds__(256) k_add_cov(DevView v, int f, int n0, int cnt)
{
    const int a = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ double Ja[kAddJ], JP[42];   // JP = J_a P77 (6 x 7)
    double* P = v.P + (size_t)f * v.nmax * v.ld;
    for (int e = threadIdx.x; e < kAddJ; e += blockDim.x) Ja[e] = v.addJ[(size_t)a * kAddJ + e];
    __syncthreads();
    if (threadIdx.x < 42) {
        const int r = threadIdx.x / 7, l = threadIdx.x % 7;
        double s = 0.0;
        for (int k = 0; k < 7; ++k) s += Ja[r * 7 + k] * P[(size_t)k * v.ld + l];
        JP[threadIdx.x] = s;
    }
    __syncthreads();
    const int row0 = n0 + 6 * a;
    if (j < n0) {
        double col[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) col[k] = P[(size_t)k * v.ld + j];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 7; ++k) s += Ja[r * 7 + k] * col[k];
            P[(size_t)(row0 + r) * v.ld + j] = s;
            P[(size_t)j * v.ld + row0 + r] = s;
        }
    } else if (j < n0 + 6 * cnt) {
        const int b = (j - n0) / 6, s = (j - n0) % 6;
        if (b > a) return;
        const double* Jb = v.addJ + (size_t)b * kAddJ;
        for (int r = 0; r < 6; ++r) {
            if (b == a && s > r) continue;
            double val = 0.0;
            for (int l = 0; l < 7; ++l) val += JP[r * 7 + l] * Jb[s * 7 + l];
            if (b == a) val += Ja[42 + r * 6 + s];
            P[(size_t)(row0 + r) * v.ld + j] = val;
            P[(size_t)j * v.ld + row0 + r] = val;
        }
    }
}

// test / host hook: one drawUncertaintyEllipse2D (Gui/Draw.cpp:42-64) into a W x H byte image, one warp
__global__ void __launch_bounds__(32) k_raster_one(uint8_t* img, int W, int H, double cx, double cy, double s00, double s01,
                                                   double s10, double s11, int maxAxes, int val)
{
    extern __shared__ __align__(16) unsigned char rs_raw[];
    RasterScratch* sc = reinterpret_cast<RasterScratch*>(rs_raw);
    int* spans = reinterpret_cast<int*>(rs_raw + sizeof(RasterScratch));
    const double S[4] = {s00, s01, s10, s11};
    float aw, ah;
    double ang;
    gate_ellipse(S, &aw, &ah, &ang);
    const float cxf = (float)cx, cyf = (float)cy;
    const float mw = fminf(aw, (float)maxAxes), mh = fminf(ah, (float)maxAxes);
    raster_ellipse_warp(img, W, H, (int)cxf, (int)cyf, (int)mw, (int)mh, ang * 180.0 / kPiTrunc, sc, spans, threadIdx.x, (uint8_t)val);
}

}  // namespace ekf
