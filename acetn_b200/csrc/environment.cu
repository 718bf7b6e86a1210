// Environment contractions behind the C ABI (SURVEY.md 8b minimum set): acetn_b200_site_rdm, acetn_b200_bond_rdm,
// acetn_b200_norm_tensor.  Reference call sites: RDM.build_site_rdm / build_bond_rdm (acetn/measurement/rdm.py:35-154) and
// build_norm_tensor (acetn/evolution/full_update.py:163-227).
//
// Host code only: every step is one K1 launch (gemm.cu) whose two-level index descriptors absorb the leg permutations of the
// reference's einsum chain, so the 2 - 8 GiB intermediates are written once in the layout the next GEMM reads; the only
// re-layouts are 64 KiB gathers of the site tensor and -- for the site RDM, whose last big contraction sums over four
// non-adjacent legs -- two gathers (the "T" of a TTGT contraction).  Nothing here synchronises or allocates.
#include "../../include/acetn_b200.h"

#include "gemm.cuh"
#include "kernels.cuh"

using namespace ab200;

namespace {

inline cudaStream_t S_(void* s) { return (cudaStream_t)s; }
inline size_t maxz(size_t a, size_t b) { return a > b ? a : b; }
inline size_t dbytes(int64_t n) { return ws_round((size_t)n * 8); }

// C[M,N] = op(A) B for row-major operands; ta: A is stored (K, M)
GemmDesc mm(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K, bool ta = false) {
    return gemm_desc((int)M, (int)N, (int)K, ta ? operand(A, idx1(1), idx1(M)) : operand(A, idx1(K), idx1(1)), operand(B, idx1(N), idx1(1)), C,
                     idx1(N), idx1(1));
}

// ---- t[(c,p,P),(e,q,Q)] = sum_{a,b} cA[a,b] eA[b,c,p,P] eB[e,a,q,Q] : the boundary part of an environment half
//      (full_update.py:205-206 / 220-221, rdm.py:95-96 / 100-101 / 56-57).  cA (x0,x1), eA (x1,xc,D,D), eB (xe,x0,D,D).
struct Front { int64_t x0, x1, xc, xe, D; };
GemmDesc front_g1(const Front& f, const double* cA, const double* eA, double* t1) {
    return mm(cA, eA, t1, f.x0, f.xc * f.D * f.D, f.x1);                                   // t1[a,(c,p,P)]
}
GemmDesc front_g2(const Front& f, const double* t1, const double* eB, double* t) {
    const int64_t D2 = f.D * f.D, m = f.xc * D2, n = f.xe * D2;
    return gemm_desc((int)m, (int)n, (int)f.x0, operand(t1, idx1(1), idx1(m)), operand(eB, idx1(D2), idx2(D2, f.x0 * D2, 1)), t, idx1(n),
                     idx1(1));
}
size_t front_gemm_ws(const Front& f) {
    return maxz(gemm_workspace_bytes(front_g1(f, nullptr, nullptr, nullptr)), gemm_workspace_bytes(front_g2(f, nullptr, nullptr, nullptr)));
}
int front(const Front& f, const double* cA, const double* eA, const double* eB, double* t1, double* t, void* g, size_t gb, cudaStream_t s) {
    AB_TRY(gemm_launch(front_g1(f, cA, eA, t1), g, gb, s));
    return gemm_launch(front_g2(f, t1, eB, t), g, gb, s);
}

// ---- closing boundary A5[f][k][(d,D)]:
//      right: tmp_r2[a,f,d,D] = cB[a,b] eC[b,f,d,D]  (full_update.py:207, rdm.py:97)   cB (xk,xb), eC (xb,xf,D,D), k = a
//      left : tmp_l2[b,f,d,D] = cB[a,b] eC[f,a,d,D]  (full_update.py:222, rdm.py:102)  cB (xa,xk), eC (xf,xa,D,D), k = b
GemmDesc closing_right_g(const double* cB, const double* eC, double* A5, int64_t xk, int64_t xb, int64_t xf, int64_t D) {
    const int64_t D2 = D * D;
    return gemm_desc((int)xk, (int)(xf * D2), (int)xb, operand(cB, idx1(xb), idx1(1)), operand(eC, idx1(xf * D2), idx1(1)), A5, idx1(D2),
                     idx2(D2, xk * D2, 1));
}
GemmDesc closing_left_g(const double* cB, const double* eC, double* A5, int64_t xa, int64_t xk, int64_t xf, int64_t D) {
    const int64_t D2 = D * D;
    return gemm_desc((int)xk, (int)(xf * D2), (int)xa, operand(cB, idx1(1), idx1(xk)), operand(eC, idx1(D2), idx2(D2, xa * D2, 1)), A5, idx1(D2),
                     idx2(D2, xk * D2, 1));
}

// ---- out[f, n-leg chi, Yb, yk] from t = front(...), the closing boundary A5[f][k-leg chi][(d,D)] and the site factors as plain
//      matrices  bra[(P,Q)][Nb],  ket[(p,q)][(d,yk)]  with Nb enumerated (D,Yb) (right half) or (Yb,D) (left half):
//        right: k-leg = c (first chi leg of t), n-leg = e        left: k-leg = e, n-leg = c
struct Back { int64_t xc, xe, xf, D, Yb, yk; bool left; };
inline int64_t back_t2_numel(const Back& b) { return b.xc * b.D * b.xe * b.D * b.D * b.Yb; }
inline int64_t back_t3_numel(const Back& b) { return b.xc * b.xe * b.D * b.D * b.Yb * b.yk; }
inline int64_t back_out_numel(const Back& b) { return b.xf * (b.left ? b.xc : b.xe) * b.Yb * b.yk; }
// 3. t2[c,p,e,q,Nb] = sum_{P,Q} t[c,p,P,e,q,Q] bra[(P,Q),Nb]
GemmDesc back_g3(const Back& b, const double* t, const double* bra, double* t2) {
    const int64_t D = b.D, D2 = D * D, n = b.xe * D2, Nb = D * b.Yb, M3 = b.xc * D * b.xe * D;
    return gemm_desc((int)M3, (int)Nb, (int)D2, operand(t, idx2(b.xe * D, D * n, D), idx2(D, n, 1)), operand(bra, idx1(Nb), idx1(1)), t2,
                     idx1(Nb), idx1(1));
}
// 4. per (c,e): t3[Nb,(d,yk)] = sum_{p,q} t2[c,p,e,q,Nb] ket[(p,q),(d,yk)], written as [k-leg chi][d][D][n-leg chi][Yb][yk]
GemmDesc back_g4(const Back& b, const double* t2, const double* ket, double* t3) {
    const int64_t D = b.D, D2 = D * D, Nb = D * b.Yb, Nk = D * b.yk;
    const int64_t nl = b.left ? b.xc : b.xe, nn = b.Yb * b.yk;
    const int64_t s_e4 = D * Nb, s_c4 = D * b.xe * D * Nb;
    const int64_t s_d = D * nl * nn, s_D = nl * nn, s_n = nn, s_Y = b.yk;
    Idx2 c_b = b.left ? idx2(b.xe, s_n, D2 * nl * nn) : idx2(b.xe, D2 * nl * nn, s_n);
    Idx2 c_m = b.left ? idx2(D, s_Y, s_D) : idx2(b.Yb, s_D, s_Y);
    GemmDesc d = gemm_desc((int)Nb, (int)Nk, (int)D2, operand(t2, idx1(1), idx2(D, b.xe * D * Nb, Nb), idx2(b.xe, s_c4, s_e4)),
                           operand(ket, idx1(Nk), idx1(1), idx1(0)), t3, c_m, idx2(b.yk, s_d, 1), 1.0, 0.0, (int)(b.xc * b.xe), c_b);
    return d;
}
// 5. out[f,(n-leg chi, Yb, yk)] = sum_{k-leg chi, d, D} A5[f,(k-leg chi, d, D)] t3[(k-leg chi, d, D), (...)]
GemmDesc back_g5(const Back& b, const double* A5, const double* t3, double* out) {
    const int64_t D2 = b.D * b.D, kl = b.left ? b.xe : b.xc, nl = b.left ? b.xc : b.xe;
    return mm(A5, t3, out, b.xf, nl * b.Yb * b.yk, kl * D2);
}
size_t back_gemm_ws(const Back& b) {
    return maxz(maxz(gemm_workspace_bytes(back_g3(b, nullptr, nullptr, nullptr)), gemm_workspace_bytes(back_g4(b, nullptr, nullptr, nullptr))),
                gemm_workspace_bytes(back_g5(b, nullptr, nullptr, nullptr)));
}
int back(const Back& b, const double* t, const double* A5, const double* bra, const double* ket, double* t2, double* t3, double* out, void* g,
         size_t gb, cudaStream_t s) {
    AB_TRY(gemm_launch(back_g3(b, t, bra, t2), g, gb, s));
    AB_TRY(gemm_launch(back_g4(b, t2, ket, t3), g, gb, s));
    return gemm_launch(back_g5(b, A5, t3, out), g, gb, s);
}

// dst (contiguous over dims[0..nd)) = src[off + sum_i idx_i * strides[i]]
int gather(double* dst, const double* src, int64_t off, int nd, const int64_t* dims, const int64_t* strides, cudaStream_t s) {
    return gather_nd_launch(dst, src + off, nd, dims, strides, s);
}

// the 20 chi extents of acetn_b200_bond_rdm / _norm_tensor, in argument order: 2 per boundary tensor (its two chi legs)
struct BondChi {
    int64_t c12[2], e12[2], e11[2], c13[2], e13[2], c21[2], e21[2], e24[2], c24[2], e23[2];
};
int read_bond_chi(const int64_t* x, BondChi* c) {
    const int64_t* p = x;
    int64_t* dst[10] = {c->c12, c->e12, c->e11, c->c13, c->e13, c->c21, c->e21, c->e24, c->c24, c->e23};
    for (int i = 0; i < 10; i++) { dst[i][0] = p[2 * i]; dst[i][1] = p[2 * i + 1]; AB_REQUIRE(dst[i][0] >= 1 && dst[i][1] >= 1, "environment: chi extents must be >= 1"); }
    // right half: c12[a,b] e12[b,c,..] e11[e,a,..] ; c13[c,b'] e13[b',f,..]      left half: c21[a,b] e21[b,c,..] e24[e,a,..] ; c24[a',e] e23[f,a',..]
    AB_REQUIRE(c->c12[1] == c->e12[0] && c->e11[1] == c->c12[0], "environment: chi legs of c12 / e12 / e11 do not match");
    AB_REQUIRE(c->c13[0] == c->e12[1] && c->c13[1] == c->e13[0], "environment: chi legs of c13 / e13 do not match e12");
    AB_REQUIRE(c->c21[1] == c->e21[0] && c->e24[1] == c->c21[0], "environment: chi legs of c21 / e21 / e24 do not match");
    AB_REQUIRE(c->c24[1] == c->e24[0] && c->c24[0] == c->e23[1], "environment: chi legs of c24 / e23 do not match e24");
    AB_REQUIRE(c->e13[1] == c->e23[0] && c->e11[0] == c->e21[1], "environment: the two halves do not share their open chi legs (f, c)");
    return OK;
}
inline Front front_right(const BondChi& c, int64_t D) { return Front{c.c12[0], c.c12[1], c.e12[1], c.e11[0], D}; }
inline Front front_left(const BondChi& c, int64_t D) { return Front{c.c21[0], c.c21[1], c.e21[1], c.e24[0], D}; }

// workspace of one half: t1, t, A5, t2, t3 (+ GEMM split-K scratch); `out` buffers are the caller's
struct HalfBufs { double *t1, *t, *A5, *t2, *t3; };
size_t half_bytes(const Front& f, const Back& b, int64_t kl) {
    const int64_t D2 = f.D * f.D;
    return dbytes(f.x0 * f.xc * D2) + dbytes(f.xc * D2 * f.xe * D2) + dbytes(b.xf * kl * D2) + dbytes(back_t2_numel(b)) + dbytes(back_t3_numel(b));
}
void half_take(Workspace& ws, const Front& f, const Back& b, int64_t kl, HalfBufs* h) {
    const int64_t D2 = f.D * f.D;
    h->t1 = ws.take<double>((size_t)(f.x0 * f.xc * D2));
    h->t = ws.take<double>((size_t)(f.xc * D2 * f.xe * D2));
    h->A5 = ws.take<double>((size_t)(b.xf * kl * D2));
    h->t2 = ws.take<double>((size_t)back_t2_numel(b));
    h->t3 = ws.take<double>((size_t)back_t3_numel(b));
}

}  // namespace

extern "C" {

// =====================================================================================================================
// norm tensor
// =====================================================================================================================
size_t acetn_b200_norm_tensor_workspace_bytes(const int64_t* chi, int64_t D, int64_t nD) {
    BondChi c;
    if (read_bond_chi(chi, &c) != OK) return 0;
    const Front fr = front_right(c, D), fl = front_left(c, D);
    const int64_t xf = c.e13[1];
    const Back br{fr.xc, fr.xe, xf, D, nD, nD, false}, bl{fl.xc, fl.xe, xf, D, nD, nD, true};
    size_t half = maxz(half_bytes(fr, br, fr.xc), half_bytes(fl, bl, fl.xe));
    size_t g = maxz(maxz(front_gemm_ws(fr), front_gemm_ws(fl)), maxz(back_gemm_ws(br), back_gemm_ws(bl)));
    g = maxz(g, gemm_workspace_bytes(gemm_desc((int)(nD * nD), (int)(nD * nD), (int)(xf * fr.xe), operand(nullptr, idx1(1), idx1(nD * nD)),
                                               operand(nullptr, idx1(nD * nD), idx1(1)), nullptr, idx1(1), idx1(1))));
    return 3 * dbytes(D * D * D * nD) + dbytes(back_out_numel(br)) + dbytes(back_out_numel(bl)) + half + g + 8192;
}

/* N12[y,x,Y,X] (nD^4) of the bond (s1, s2, k); argument names follow full_update.py:163-227:
 *   right half (site s1): c12 = C[(k+1)%4], e12 = E[(k+1)%4], e11 = E[k], c13 = C[(k+2)%4], e13 = E[(k+2)%4], a1q (D,D,D,nD)
 *   left  half (site s2): c21 = C[k], e21 = E[k], e24 = E[(k+3)%4], c24 = C[(k+3)%4], e23 = E[(k+2)%4],     a2q (D,D,D,nD)
 *   n1[f,e,Y,y] = tmp_r2[a,f,d,D] ( ((c12 e12) e11) conj(a1q)[R,D,U,Y] a1q[r,d,u,y] ),  n2[f,c,X,x] likewise (left),
 *   N12 = sum_{f,c} n1[f,c,Y,y] n2[f,c,X,x]. */
int acetn_b200_norm_tensor(const double* c12, const double* e12, const double* e11, const double* c13, const double* e13, const double* a1q,
                           const double* c21, const double* e21, const double* e24, const double* c24, const double* e23, const double* a2q,
                           const int64_t* chi, int64_t D, int64_t nD, double* n12, void* wsp, size_t ws_bytes, void* stream) {
    cudaStream_t s = S_(stream);
    BondChi c;
    AB_TRY(read_bond_chi(chi, &c));
    AB_REQUIRE(D >= 1 && nD >= 1, "norm_tensor: D, nD must be >= 1");
    const Front fr = front_right(c, D), fl = front_left(c, D);
    const int64_t xf = c.e13[1], D2 = D * D;
    const Back br{fr.xc, fr.xe, xf, D, nD, nD, false}, bl{fl.xc, fl.xe, xf, D, nD, nD, true};
    Workspace ws(wsp, ws_bytes);
    double* q1 = ws.take<double>((size_t)(D2 * D * nD));
    double* bra2 = ws.take<double>((size_t)(D2 * D * nD));
    double* ket2 = ws.take<double>((size_t)(D2 * D * nD));
    double* n1 = ws.take<double>((size_t)back_out_numel(br));
    double* n2 = ws.take<double>((size_t)back_out_numel(bl));
    const size_t mark = ws.used;
    HalfBufs h;
    half_take(ws, fr, br, fr.xc, &h);
    size_t used_r = ws.used;
    ws.used = mark;
    HalfBufs hl;
    half_take(ws, fl, bl, fl.xe, &hl);
    if (used_r > ws.used) ws.used = used_r;
    if (ws.overflow || ws.used > ws.bytes) { set_error("norm_tensor: workspace too small (%zu needed, %zu given)", ws.used, ws_bytes); return ERR_WORKSPACE; }
    void* g = ws.base + ws.used;
    size_t gb = ws.bytes - ws.used;
    // the 64 KiB site factors as plain (k, n) matrices (conj is the identity: FP64 real); a?q strides: (D^2 nD, D nD, nD, 1)
    const int64_t s0 = D2 * nD, s1 = D * nD, s2 = nD, s3 = 1;
    {   // q1[(R,U)][(D,Y)] = a1q[R,D,U,Y]  (bra and ket of the right half)
        int64_t dims[4] = {D, D, D, nD}, st[4] = {s0, s2, s1, s3};
        AB_TRY(gather(q1, a1q, 0, 4, dims, st, s));
    }
    {   // bra2[(U,L)][(X,D)] = a2q[D,L,U,X]
        int64_t dims[4] = {D, D, nD, D}, st[4] = {s2, s1, s3, s0};
        AB_TRY(gather(bra2, a2q, 0, 4, dims, st, s));
    }
    {   // ket2[(u,l)][(d,x)] = a2q[d,l,u,x]
        int64_t dims[4] = {D, D, D, nD}, st[4] = {s2, s1, s0, s3};
        AB_TRY(gather(ket2, a2q, 0, 4, dims, st, s));
    }
    AB_TRY(front(fr, c12, e12, e11, h.t1, h.t, g, gb, s));
    AB_TRY(gemm_launch(closing_right_g(c13, e13, h.A5, fr.xc, c.c13[1], xf, D), g, gb, s));
    AB_TRY(back(br, h.t, h.A5, q1, q1, h.t2, h.t3, n1, g, gb, s));                              // n1[f,e,Y,y]
    AB_TRY(front(fl, c21, e21, e24, hl.t1, hl.t, g, gb, s));
    AB_TRY(gemm_launch(closing_left_g(c24, e23, hl.A5, c.c24[0], fl.xe, xf, D), g, gb, s));
    AB_TRY(back(bl, hl.t, hl.A5, bra2, ket2, hl.t2, hl.t3, n2, g, gb, s));                      // n2[f,c,X,x]
    // N12[y,x,Y,X] = sum_{(f,c)} n1[(f,c),(Y,y)] n2[(f,c),(X,x)] : m = (Y,y), n = (X,x), written straight into the [y,x,Y,X] layout
    const int64_t nn = nD * nD;
    return gemm_launch(gemm_desc((int)nn, (int)nn, (int)(xf * fr.xe), operand(n1, idx1(1), idx1(nn)), operand(n2, idx1(nn), idx1(1)), n12,
                                 idx2(nD, nD, nn * nD), idx2(nD, 1, nn)), g, gb, s);
}

// =====================================================================================================================
// bond RDM
// =====================================================================================================================
size_t acetn_b200_bond_rdm_workspace_bytes(const int64_t* chi, int64_t D, int64_t d) {
    BondChi c;
    if (read_bond_chi(chi, &c) != OK) return 0;
    const Front fr = front_right(c, D), fl = front_left(c, D);
    const int64_t xf = c.e13[1], D2 = D * D;
    const Back br{fr.xc, fr.xe, xf, D, D, D * d, false}, bl{fl.xc, fl.xe, xf, D, D, D * d, true};
    size_t half = maxz(half_bytes(fr, br, fr.xc), half_bytes(fl, bl, fl.xe));
    size_t g = maxz(maxz(front_gemm_ws(fr), front_gemm_ws(fl)), maxz(back_gemm_ws(br), back_gemm_ws(bl)));
    g = maxz(g, gemm_workspace_bytes(gemm_desc((int)d, (int)d, (int)(xf * fr.xe * D2), operand(nullptr, idx1(1), idx1(d)),
                                               operand(nullptr, idx1(d), idx1(1)), nullptr, idx1(d), idx1(1))));
    return 2 * dbytes(D2 * D2 * d) + 2 * dbytes(D2 * D2) + (size_t)d * (dbytes(back_out_numel(br)) + dbytes(back_out_numel(bl))) + half + g + 8192;
}

/* rho[P,Q,p,q] (d^4) of the bond (s1, s2, k) (rdm.py:69-154, blocked over the bra physical index like build_bond_rdm_core_blocked,
 * each half computed d times instead of d^2 times).  Boundary tensors as acetn_b200_norm_tensor; a1 / a2 = site.bond_permute(k) of s1 /
 * s2: strided views (D,D,D,D,d) with 5 element strides each.
 *   right[P][f,c,L,(l,p)] = tmp_r2[a,f,d,D] ( tr1 conj(a1)[L,U,R,D,P] a1[l,u,r,d,p] )      (rdm.py:95-99, 133-139)
 *   left [Q][f,c,R,(r,q)] = tmp_l2[e,f,d,D] ( tl1 conj(a2)[L,U,R,D,Q] a2[l,u,r,d,q] )      (rdm.py:100-104, 141-147)
 *   rho[P,Q,p,q] = sum right[P][f,c,R,r,p] left[Q][f,c,R,r,q]. */
int acetn_b200_bond_rdm(const double* c12, const double* e12, const double* e11, const double* c13, const double* e13, const double* a1,
                        const int64_t* a1_strides, const double* c21, const double* e21, const double* e24, const double* c24,
                        const double* e23, const double* a2, const int64_t* a2_strides, const int64_t* chi, int64_t D, int64_t d, double* rho,
                        void* wsp, size_t ws_bytes, void* stream) {
    cudaStream_t s = S_(stream);
    BondChi c;
    AB_TRY(read_bond_chi(chi, &c));
    AB_REQUIRE(D >= 1 && d >= 1, "bond_rdm: D, d must be >= 1");
    const Front fr = front_right(c, D), fl = front_left(c, D);
    const int64_t xf = c.e13[1], D2 = D * D;
    const Back br{fr.xc, fr.xe, xf, D, D, D * d, false}, bl{fl.xc, fl.xe, xf, D, D, D * d, true};
    Workspace ws(wsp, ws_bytes);
    double* ket1 = ws.take<double>((size_t)(D2 * D2 * d));
    double* ket2 = ws.take<double>((size_t)(D2 * D2 * d));
    double* bra = ws.take<double>((size_t)(D2 * D2));
    double* bra_b = ws.take<double>((size_t)(D2 * D2));
    double* right[16];
    double* left[16];
    AB_REQUIRE(d <= 16, "bond_rdm: physical dimension above 16 is not supported");
    for (int P = 0; P < d; P++) right[P] = ws.take<double>((size_t)back_out_numel(br));
    for (int Q = 0; Q < d; Q++) left[Q] = ws.take<double>((size_t)back_out_numel(bl));
    const size_t mark = ws.used;
    HalfBufs h;
    half_take(ws, fr, br, fr.xc, &h);
    size_t used_r = ws.used;
    ws.used = mark;
    HalfBufs hl;
    half_take(ws, fl, bl, fl.xe, &hl);
    if (used_r > ws.used) ws.used = used_r;
    if (ws.overflow || ws.used > ws.bytes) { set_error("bond_rdm: workspace too small (%zu needed, %zu given)", ws.used, ws_bytes); return ERR_WORKSPACE; }
    void* g = ws.base + ws.used;
    size_t gb = ws.bytes - ws.used;
    const int64_t* as = a1_strides;
    const int64_t* bs = a2_strides;
    {   // ket1[(r,u)][(d,(l,p))] = a1[l,u,r,d,p]
        int64_t dims[5] = {D, D, D, D, d}, st[5] = {as[2], as[1], as[3], as[0], as[4]};
        AB_TRY(gather(ket1, a1, 0, 5, dims, st, s));
    }
    {   // ket2[(u,l)][(d,(r,q))] = a2[l,u,r,d,q]
        int64_t dims[5] = {D, D, D, D, d}, st[5] = {bs[1], bs[0], bs[3], bs[2], bs[4]};
        AB_TRY(gather(ket2, a2, 0, 5, dims, st, s));
    }
    AB_TRY(front(fr, c12, e12, e11, h.t1, h.t, g, gb, s));
    AB_TRY(gemm_launch(closing_right_g(c13, e13, h.A5, fr.xc, c.c13[1], xf, D), g, gb, s));
    for (int P = 0; P < d; P++) {
        int64_t dims[4] = {D, D, D, D}, st[4] = {as[2], as[1], as[3], as[0]};                  // bra[(R,U)][(D,L)] = a1[L,U,R,D,P]
        AB_TRY(gather(bra, a1, (int64_t)P * as[4], 4, dims, st, s));
        AB_TRY(back(br, h.t, h.A5, bra, ket1, h.t2, h.t3, right[P], g, gb, s));
    }
    AB_TRY(front(fl, c21, e21, e24, hl.t1, hl.t, g, gb, s));
    AB_TRY(gemm_launch(closing_left_g(c24, e23, hl.A5, c.c24[0], fl.xe, xf, D), g, gb, s));
    for (int Q = 0; Q < d; Q++) {
        int64_t dims[4] = {D, D, D, D}, st[4] = {bs[1], bs[0], bs[2], bs[3]};                  // bra[(U,L)][(R,D)] = a2[L,U,R,D,Q]
        AB_TRY(gather(bra_b, a2, (int64_t)Q * bs[4], 4, dims, st, s));
        AB_TRY(back(bl, hl.t, hl.A5, bra_b, ket2, hl.t2, hl.t3, left[Q], g, gb, s));
    }
    const int64_t K = xf * fr.xe * D2;
    for (int P = 0; P < d; P++)
        for (int Q = 0; Q < d; Q++)
            AB_TRY(gemm_launch(gemm_desc((int)d, (int)d, (int)K, operand(right[P], idx1(1), idx1(d)), operand(left[Q], idx1(d), idx1(1)),
                                         rho + ((int64_t)P * d + Q) * d * d, idx1(d), idx1(1)), g, gb, s));
    return OK;
}

// =====================================================================================================================
// site RDM
// =====================================================================================================================
namespace {
struct SiteChi { int64_t c1[2], c2[2], c3[2], c4[2], e1[2], e2[2], e3[2], e4[2]; };
int read_site_chi(const int64_t* x, SiteChi* c) {
    int64_t* dst[8] = {c->c1, c->c2, c->c3, c->c4, c->e1, c->e2, c->e3, c->e4};
    for (int i = 0; i < 8; i++) { dst[i][0] = x[2 * i]; dst[i][1] = x[2 * i + 1]; AB_REQUIRE(dst[i][0] >= 1 && dst[i][1] >= 1, "site_rdm: chi extents must be >= 1"); }
    // rdm.py:56-65: c4[a,b] e4[b,c,..] e3[e,a,..] ; c1[a,b] e1[b,c,..] ; c3[a,b] e2[c,a,..] ; c2[e,c] ; closed ring
    AB_REQUIRE(c->c4[1] == c->e4[0] && c->e3[1] == c->c4[0], "site_rdm: chi legs of c4 / e4 / e3 do not match");
    AB_REQUIRE(c->c1[1] == c->e1[0], "site_rdm: chi legs of c1 / e1 do not match");
    AB_REQUIRE(c->e2[1] == c->c3[0] && c->c2[1] == c->e2[0], "site_rdm: chi legs of c3 / e2 / c2 do not match");
    AB_REQUIRE(c->c2[0] == c->e1[1], "site_rdm: chi legs of c2 / e1 do not match");
    AB_REQUIRE(c->c3[1] == c->e3[0] && c->c1[0] == c->e4[1], "site_rdm: the corner ring does not close (c3 / e3, c1 / e4)");
    return OK;
}
struct SiteBufs {
    double *bra, *ket, *t1a, *t1, *t2, *tB1, *tB2, *tB3, *tB4, *G1, *G2, *t3f;
};
struct SiteDims { int64_t x0, x1, xc, xe, xa, xk1, xe1, xb, xk3, xc3, D, d; };
inline SiteDims site_dims(const SiteChi& c, int64_t D, int64_t d) {
    SiteDims q;
    q.x0 = c.c4[0]; q.x1 = c.c4[1]; q.xc = c.e4[1]; q.xe = c.e3[0];      // ((c4 e4) e3): open legs c (-> c1 rows), e (-> c3 cols)
    q.xa = c.c1[0]; q.xk1 = c.c1[1]; q.xe1 = c.e1[1];                     // tB1[a,e',uU]: a = c1 rows (= xc), e' = e1 cols (= c2 rows)
    q.xb = c.c3[1]; q.xk3 = c.c3[0]; q.xc3 = c.e2[0];                     // tB2[b,c,rR]: b = c3 cols (= xe), c = e2 rows (= c2 cols)
    q.D = D; q.d = d;
    return q;
}
// the eight GEMMs of rdm.py:56-66 after the boundary front, in launch order
int site_gemms(const SiteDims& q, const double* c1, const double* c2, const double* c3, const double* e1, const double* e2, const SiteBufs& b,
               double* rho, GemmDesc out[8]) {
    const int64_t D = q.D, D2 = D * D, Nb = D2 * q.d, n = q.xe * D2, M3 = q.xc * D * q.xe * D;
    // t2[(c,l,e,d),(U,R,P)] = sum_{L,D} t1[c,l,L,e,d,D] bra[(L,D),(U,R,P)]
    out[0] = gemm_desc((int)M3, (int)Nb, (int)D2, operand(b.t1, idx2(q.xe * D, D * n, D), idx2(D, n, 1)), operand(b.bra, idx1(Nb), idx1(1)), b.t2,
                       idx1(Nb), idx1(1));
    // tmp2 = c1 e1: tB1[a,(e',u,U)]
    out[1] = mm(c1, e1, b.tB1, q.xa, q.xe1 * D2, q.xk1);
    // tmp3 = c3 e2: tB2[b,c,(r,R)] = sum_a c3[a,b] e2[c,a,(r,R)]
    out[2] = gemm_desc((int)q.xb, (int)(q.xc3 * D2), (int)q.xk3, operand(c3, idx1(1), idx1(q.xb)), operand(e2, idx1(D2), idx2(D2, q.xk3 * D2, 1)),
                       b.tB2, idx1(q.xc3 * D2), idx1(1));
    // tB3[e',b,(r,R)] = sum_c c2[e',c] tB2[b,c,(r,R)]
    out[3] = gemm_desc((int)q.xe1, (int)(q.xb * D2), (int)q.xc3, operand(c2, idx1(q.xc3), idx1(1)), operand(b.tB2, idx1(D2), idx2(D2, q.xc3 * D2, 1)),
                       b.tB3, idx1(q.xb * D2), idx1(1));
    // tB4[(b,r,R),(a,u,U)] = sum_e' tB3[e',(b,r,R)] tB1[a,e',(u,U)]
    out[4] = gemm_desc((int)(q.xb * D2), (int)(q.xa * D2), (int)q.xe1, operand(b.tB3, idx1(1), idx1(q.xb * D2)),
                       operand(b.tB1, idx1(D2), idx2(D2, q.xe1 * D2, 1)), b.tB4, idx1(q.xa * D2), idx1(1));
    // t3f[(r,u),(l,d,P)] = G1[(r,u),(e,c,R,U)] G2[(e,c,R,U),(l,d,P)]
    out[5] = mm(b.G1, b.G2, b.t3f, D2, D2 * q.d, q.xb * q.xa * D2);
    // rho[P,p] = sum_{r,u,l,d} t3f[(r,u,l,d),P] ket[(r,u,l,d),p]
    out[6] = gemm_desc((int)q.d, (int)q.d, (int)(D2 * D2), operand(b.t3f, idx1(1), idx1(q.d)), operand(b.ket, idx1(q.d), idx1(1)), rho, idx1(q.d),
                       idx1(1));
    return 7;
}
void site_take(Workspace& ws, const SiteDims& q, SiteBufs* b) {
    const int64_t D = q.D, D2 = D * D, Nb = D2 * q.d;
    b->bra = ws.take<double>((size_t)(D2 * Nb));
    b->ket = ws.take<double>((size_t)(D2 * D2 * q.d));
    b->t1a = ws.take<double>((size_t)(q.x0 * q.xc * D2));
    b->t1 = ws.take<double>((size_t)(q.xc * D2 * q.xe * D2));
    b->t2 = ws.take<double>((size_t)(q.xc * D * q.xe * D * Nb));
    b->tB1 = ws.take<double>((size_t)(q.xa * q.xe1 * D2));
    b->tB2 = ws.take<double>((size_t)(q.xb * q.xc3 * D2));
    b->tB3 = ws.take<double>((size_t)(q.xe1 * q.xb * D2));
    b->tB4 = ws.take<double>((size_t)(q.xb * D2 * q.xa * D2));
    b->G1 = ws.take<double>((size_t)(q.xb * D2 * q.xa * D2));
    b->G2 = ws.take<double>((size_t)(q.xc * D * q.xe * D * Nb));
    b->t3f = ws.take<double>((size_t)(D2 * D2 * q.d));
}
}  // namespace

size_t acetn_b200_site_rdm_workspace_bytes(const int64_t* chi, int64_t D, int64_t d) {
    SiteChi c;
    if (read_site_chi(chi, &c) != OK) return 0;
    const SiteDims q = site_dims(c, D, d);
    Workspace ws(nullptr, 0);                    // dry run: only `used` is of interest
    SiteBufs b;
    site_take(ws, q, &b);
    GemmDesc gd[8];
    const int ng = site_gemms(q, nullptr, nullptr, nullptr, nullptr, nullptr, b, nullptr, gd);
    size_t g = front_gemm_ws(Front{q.x0, q.x1, q.xc, q.xe, D});
    for (int i = 0; i < ng; i++) g = maxz(g, gemm_workspace_bytes(gd[i]));
    return ws.used + g + 8192;
}

/* rho[P,p] (bra, ket) of one site (rdm.py:35-67).  c1..c4 = site.C[0..3], e1..e4 = site.E[0..3], A = site['A'] (D,D,D,D,d) with 5 element
 * strides.  chi: 16 extents, 2 per tensor in the order c1, c2, c3, c4, e1, e2, e3, e4. */
int acetn_b200_site_rdm(const double* c1, const double* c2, const double* c3, const double* c4, const double* e1, const double* e2,
                        const double* e3, const double* e4, const double* A, const int64_t* a_strides, const int64_t* chi, int64_t D, int64_t d,
                        double* rho, void* wsp, size_t ws_bytes, void* stream) {
    cudaStream_t s = S_(stream);
    SiteChi c;
    AB_TRY(read_site_chi(chi, &c));
    AB_REQUIRE(D >= 1 && d >= 1, "site_rdm: D, d must be >= 1");
    const SiteDims q = site_dims(c, D, d);
    const int64_t D2 = D * D, Nb = D2 * d;
    Workspace ws(wsp, ws_bytes);
    SiteBufs b;
    site_take(ws, q, &b);
    if (ws.overflow) { set_error("site_rdm: workspace too small (%zu needed, %zu given)", ws.used, ws_bytes); return ERR_WORKSPACE; }
    void* g = ws.base + ws.used;
    size_t gb = ws.bytes - ws.used;
    const int64_t* as = a_strides;
    {   // bra[(L,D)][(U,R,P)] = A[L,U,R,D,P]   (conj: identity)
        int64_t dims[5] = {D, D, D, D, d}, st[5] = {as[0], as[3], as[1], as[2], as[4]};
        AB_TRY(gather(b.bra, A, 0, 5, dims, st, s));
    }
    {   // ket[(r,u,l,d)][p] = A[l,u,r,d,p]
        int64_t dims[5] = {D, D, D, D, d}, st[5] = {as[2], as[1], as[0], as[3], as[4]};
        AB_TRY(gather(b.ket, A, 0, 5, dims, st, s));
    }
    GemmDesc gd[8];
    site_gemms(q, c1, c2, c3, e1, e2, b, rho, gd);
    // tmp1 = ((c4 e4) e3) conj(a1)
    AB_TRY(front(Front{q.x0, q.x1, q.xc, q.xe, D}, c4, e4, e3, b.t1a, b.t1, g, gb, s));        // t1[(c,l,L),(e,d,D)]
    for (int i = 0; i < 5; i++) AB_TRY(gemm_launch(gd[i], g, gb, s));
    // "erRcuU,cledURP->ruldP": tB4 relabelled [e,r,R,c,u,U] (e = b, c = a), t2 [c,l,e,d,U,R,P]; the contracted legs (e,c,R,U) are not
    // adjacent in either operand, so both are gathered once (the transposes of a TTGT contraction)
    {   // G1[(r,u)][(e,c,R,U)]
        int64_t dims[6] = {D, D, q.xb, q.xa, D, D};
        int64_t st[6] = {D * q.xa * D2, D, D2 * q.xa * D2, D2, q.xa * D2, 1};
        AB_TRY(gather(b.G1, b.tB4, 0, 6, dims, st, s));
    }
    {   // G2[(e,c,R,U)][(l,d,P)] from t2[c,l,e,d,U,R,P]
        const int64_t sP = 1, sR = d, sU = D * d, sd = Nb, se = D * Nb, sl = q.xe * D * Nb, sc = D * q.xe * D * Nb;
        int64_t dims[7] = {q.xe, q.xc, D, D, D, D, d};
        int64_t st[7] = {se, sc, sR, sU, sl, sd, sP};
        AB_TRY(gather(b.G2, b.t2, 0, 7, dims, st, s));
    }
    AB_TRY(gemm_launch(gd[5], g, gb, s));
    return gemm_launch(gd[6], g, gb, s);
}

}  // extern "C"
