// K3: the dense factorisation (probdiffeq/_probdiffeq/ssm_impl_dense.py), one IVP instance per CTA, COLUMNS IN REGISTERS.
//
// The state is a mean of length N = n d (coefficient-major, index k d + i) and a full N x N left square root.
// Every step triangularises a 2N x N stack (extrapolation, DenseLatentCond.marginalise :24-33) and an
// (N + d) x (N + d) stack (correction, DenseLatentCond.revert :51-77 with util/cholesky_util.py:27-82); for
// HIRES (d = 8, nu = 5) these are 96 x 48 and 56 x 56, ~110 dependent Householder reflectors per attempt.
//
// Mapping. A column of the stack lives in the REGISTERS of TPC (2 or 4) adjacent lanes, rows dealt round-robin
// (row r -> lane part r % TPC, register r / TPC); the CTA holds N "state" columns and d "observation" columns.
// One reflector step: the lanes that own the pivot column publish it RAW -- its part below the pivot, zero outside its
// row extent so that nobody needs a run-time register index, and the pivot element apart -- in one of two
// shared-memory buffers; ONE block barrier; every lane then reads its row part of that buffer once and forms, from the
// same numbers in the same order as every other lane, the column's norm, the reflector scalars (LAPACK dlarfg
// convention, see pdeq_blockops.cuh) and its part of v^T c, meets its partners with a shuffle and updates its
// registers. Nobody waits for an owner: round 2's first version had the owning warp reduce the norm, form the
// reflector and publish it (and clear its registers) -- ~295 instructions of one warp per pivot, with the other three
// at the barrier, against ~85 for the update; now a pivot costs ~45 (publish) + ~145 (everything else) instructions
// of the busiest warp (ncu: config 4a 3.57 -> 3.10 s; what remains is the issue rate of ONE warp per instance,
// ~5 cycles per instruction whatever the instruction, with three instances per SM). The trailing matrix is never read or written in shared
// memory (round 1's kernel did both for every column and was bound by exactly that: ~2600 cycles per column with
// four instances per SM), the reflector loop is ROLLED (a few hundred instructions, so the kernel lives in the
// instruction cache), and the shared memory per instance drops from 55 KB to the packed accepted factor plus vectors.
// Stack rows are ordered [(H L)^T, L^T ; damp I, 0] -- a row permutation of the reference's block matrix that leaves
// R^T R, hence all covariances and the gain, unchanged -- so that the extrapolated factor is reused in place: the
// registers that held column c of the extrapolation stack hold column d + c of the correction stack.
//
// Restates for the dense model: DenseWienerIntegrated.transition (:347-363), DenseOdeTs0.linearize (:243-259),
// DenseResidual.linearize (:290-334, exact Jacobian, jacobians.py:93-98), DenseNormal.std (:144-150),
// residual_whitened_rms_flat (:156-160); the loop / solver / error / control lines are those listed in
// pdeq_loop_thread.cuh.  Filter strategy only.
#pragma once

#include "pdeq_limits.cuh"
#include "pdeq_loop_thread.cuh"

namespace pdeq {

template <class VF, int NU, bool TS0>
struct DenseLoop {
  static constexpr int n = NU + 1;
  static constexpr int q = VF::order;
  static constexpr int P = VF::num_params > 0 ? VF::num_params : 1;
  // All extents are compile-time: the dense kernels exist for fixed-dimension vector fields only (the host checks
  // cfg.ode_dim == VF::fixed_dim), so every index split (e / N, e % d, ...) is a multiply-shift, not a division.
  static constexpr int D = VF::fixed_dim;
  static_assert(D > 0, "the dense kernels need a compile-time ODE dimension");
  static constexpr int N = n * D, LD = N + D, HW = (q + 1) * D, TRI_N = N * (N + 1) / 2;
  static constexpr int TPC = DenseSmemLayout::tpc(N);               // lanes per column
  static constexpr int RPT = DenseSmemLayout::rows_per_thread(N);   // registers per lane (rows of the 2N-row stack)
  static constexpr int RPAD = DenseSmemLayout::rpad(N);
  static constexpr int RB = TPC * RPAD + 2;                         // one reflector buffer: TPC row parts, tp
  static constexpr int G = DenseSmemLayout::threads(N, D);
  static_assert(TPC * RPT >= 2 * N, "rows do not fit");

  struct VecAcc {
    const double* u;  // coefficient-major mean
    int d;
    PDEQ_DI double operator()(int k, int i) const { return u[k * d + i]; }
  };

  PDEQ_DI static int tri(int i, int j) { return i * (i + 1) / 2 + j; }  // packed lower, j <= i

  // sum over the TPC lanes that share a column
  PDEQ_DI static double group_sum(double v, unsigned gmask) {
#pragma unroll
    for (int o = 1; o < TPC; o <<= 1) v += __shfl_xor_sync(gmask, v, o);
    return v;
  }

  PDEQ_DI static double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < G / 32; ++w) tot += red[w];
    return tot;
  }

  // Which rows a reflector touches. Pivot j (row j) reaches the rows r > j of the top block (r < SA) and, of the bottom
  // block, row SA + k once k <= j (the bottom block of both stacks is upper triangular: row k enters with pivot k) --
  // all of them for j >= CSPLIT (fill-in has made the trailing matrix dense by then).
  PDEQ_DI static bool row_in(int r, int j, int csplit, int sa) {
    return r > j && (r < sa || j >= csplit || r - sa <= j);
  }
  // The same for a BLOCK of pivots J0 <= j < J1, decided at compile time: 0 = no pivot of the block touches row r,
  // 1 = every pivot does (and r is not a pivot row), 2 = it depends on j. A register holds TPC consecutive rows (which
  // of them is this lane's is a run-time matter), so registers are classified by the union over their rows.
  static constexpr int row_class(int r, int J0, int J1, int M, int CSPLIT, int SA) {
    if (r >= M || r < J0) return 0;
    if (r < J1) return 2;
    if (r < SA || J0 >= CSPLIT) return 1;
    const int k = r - SA;
    if (k < J0) return 1;
    if (J1 <= CSPLIT && k >= J1) return 0;
    return 2;
  }
  static constexpr int reg_class(int i, int J0, int J1, int M, int CSPLIT, int SA) {
    int c = row_class(i * TPC, J0, J1, M, CSPLIT, SA);
    for (int hh = 1; hh < TPC; ++hh) {
      const int c2 = row_class(i * TPC + hh, J0, J1, M, CSPLIT, SA);
      if (c2 != c) c = 2;
    }
    return c;
  }

  // Householder triangularisation of the M-row stack whose columns live in the lanes' registers.
  // `mypos`: position of this lane's column in the elimination order (-1: no column); `first_lane(j)`: the first lane
  // of the group that owns pivot j. After the call a column holds R above and on its diagonal and exact zeros below.
  //
  // One block barrier per reflector (the two reflector buffers alternate: whoever writes buffer (j & 1) again has
  // passed barrier j + 1, which every reader of step j reaches only after its reads). The pivots are taken in blocks
  // of eight, unrolled, so that inside a block the rows a reflector can touch are known at compile time (row_class):
  // a step costs what its rows cost, not what the whole column costs, with run-time tests only for the eight rows
  // around the pivots. Nothing in a step diverges inside a warp except the owner's stores: every lane forms the
  // reflector, a lane whose column is finished updates it with weight zero (shuffles under a diverged mask go through
  // a slow path that cost more than the arithmetic they saved), and the eliminated entries of the pivot columns stay
  // where they are until one clearing pass at the end.
  template <int M, int NPOS, int CSPLIT, int SA, class FirstLane>
  PDEQ_DI static void column_qr(double (&col)[RPT], int mypos, int h, FirstLane first_lane, double* rbuf) {
    constexpr int RM = (M + TPC - 1) / TPC;  // registers in use for an M-row stack
    constexpr int BS = 8;
    const int warp = threadIdx.x >> 5;
    static_for<0, (NPOS + BS - 1) / BS>([&](auto jb_) {
      constexpr int J0 = decltype(jb_)::value * BS;
      constexpr int J1 = imin(J0 + BS, NPOS);
      for (int j = J0; j < J1; ++j) {
        double* buf = rbuf + (j & 1) * RB;
        if (mypos == j) {  // publish the RAW pivot column: its part below the pivot, and the pivot element apart
          static_for<0, RM>([&](auto ic_) {
            constexpr int i = decltype(ic_)::value;
            constexpr int cls = reg_class(i, J0, J1, M, CSPLIT, SA);
            if constexpr (cls == 1) {
              buf[h * RPAD + i] = col[i];
            } else if constexpr (cls == 2) {
              const int r = i * TPC + h;
              buf[h * RPAD + i] = row_in(r, j, CSPLIT, SA) ? col[i] : 0.0;
              if (r == j) buf[TPC * RPAD] = col[i];
            }
          });
        }
        __syncthreads();
        if (__any_sync(0xffffffffu, mypos >= j)) {  // warp-uniform: this warp still holds the pivot or a trailing column
          const double* v = buf + h * RPAD;
          const double alpha = buf[TPC * RPAD];
          double w0 = 0.0, w1 = 0.0, ss0 = 0.0, ss1 = 0.0, cj = 0.0;
          static_for<0, RM>([&](auto ic_) {
            constexpr int i = decltype(ic_)::value;
            constexpr int cls = reg_class(i, J0, J1, M, CSPLIT, SA);
            if constexpr (cls != 0) {
              const double x = v[i];
              if constexpr (i & 1) w1 = fma(x, col[i], w1);
              else w0 = fma(x, col[i], w0);
              if constexpr (cls == 1 && (i & 1)) ss1 = fma(x, x, ss1);
              else ss0 = fma(x, x, ss0);
              if constexpr (cls == 2) cj = (i * TPC + h == j) ? col[i] : cj;
            }
          });
          // every lane forms the reflector from the published column -- the same arithmetic on the same numbers in
          // every lane, so nobody waits for an owner: dlarfg, as in the owner-computes scheme
          const double ss = group_sum(ss0 + ss1, 0xffffffffu);
          const double wd = group_sum(w0 + w1, 0xffffffffu);
          cj = group_sum(cj, 0xffffffffu);
          const bool live = ss != 0.0;  // dlarfg: xnorm == 0 -> tau = 0, H = I
          const double tt = fma(alpha, alpha, ss);
          const double y = fast_rsqrt(live ? tt : 1.0);
          const double nrm = tt * y;
          const double sgn_nrm = copysign(nrm, alpha);
          const double v0 = alpha + sgn_nrm;
          const double tp = live ? fast_rcp(fma(nrm, fabs(alpha), tt)) : 0.0;
          const double beta = live ? -sgn_nrm : alpha;
          const double w = (mypos > j) ? fma(v0, cj, wd) * tp : 0.0;
          const double wv0 = w * v0;
          const bool own = mypos == j;
          static_for<0, RM>([&](auto ic_) {
            constexpr int i = decltype(ic_)::value;
            constexpr int cls = reg_class(i, J0, J1, M, CSPLIT, SA);
            if constexpr (cls == 1) {
              col[i] = fma(-w, v[i], col[i]);
            } else if constexpr (cls == 2) {
              const double upd = fma(-w, v[i], col[i]);
              col[i] = (i * TPC + h == j) ? (own ? beta : col[i] - wv0) : upd;
            }
          });
        }
      }
    });
    // the eliminated part of every pivot column, left in place by the steps above (later reflectors pass over a
    // finished column with weight zero), is cleared once
    static_for<0, RM>([&](auto ic_) {
      constexpr int i = decltype(ic_)::value;
      col[i] = (mypos >= 0 && mypos < NPOS && i * TPC + h > mypos) ? 0.0 : col[i];
    });
    __syncthreads();  // the last buffer may be rewritten by whatever comes next
  }

  // Write one checkpoint (mean [n][d], chol [N][N]) from shared memory.
  PDEQ_DI static void emit(const LoopArgs& a, long b, int ck, double t, const double* m, const double* Lpk,
                           double chol_scale, double scale, int nsteps) {
    const long bt = b * a.T + ck;
    if (threadIdx.x == 0) {
      a.sol.t[bt] = t;
      a.sol.num_steps[bt] = nsteps;
      if (a.sol.output_scale != nullptr) a.sol.output_scale[bt] = scale;
    }
    for (int e = threadIdx.x; e < N; e += G) a.sol.mean[bt * N + e] = m[e];
    if (a.sol.chol != nullptr) {
      double* co = a.sol.chol + bt * (long)N * N;
      for (int e = threadIdx.x; e < N * N; e += G) {
        const int i = e / N, j = e % N;
        co[e] = (j <= i) ? chol_scale * Lpk[tri(i, j)] : 0.0;
      }
    }
  }

  // Everything the per-step routines share: this lane's place in the column layout and the instance's shared memory.
  // (Plain forced-inline functions taking the column registers by reference: a lambda that captures `col` and is not
  // inlined takes its address and puts all of it in local memory -- measured: 4x slower than round 1's kernel.)
  struct Ctx {
    int tid, cp, h, ci, cj, oa, pos_pred, pos_obs, pos_rev;
    bool is_state, is_obs;
    unsigned gmask;
    double inv_sqrt_d;
    const double (*A)[PDEQ_MAX_COEFFS];
    const double (*Qm)[PDEQ_MAX_COEFFS];
    double *rbuf, *Hs, *Ush, *RY, *mobs, *wht, *stdv, *lam, *p, *pinv;
  };

  // Linearise the constraint at `mean` (DenseOdeTs0.linearize :243-259 / DenseResidual.linearize :290-334): the
  // non-zero columns of H go to Hs, the observed mean H m + bias to mobs. Ends with a block barrier.
  PDEQ_DI static void linearise_at(const Ctx& c, const double* mean, double tt, const double (&params)[P]) {
    constexpr int d = D, hw = HW;
    for (int e = c.tid; e < d * hw; e += G) c.Hs[e] = 0.0;
    __syncthreads();
    VecAcc acc{mean, d};
    for (int jd = c.tid; jd < d; jd += G) {
      const double f = VF::template component<double>(jd, d, acc, params, tt);
      c.Hs[jd * hw + q * d + jd] = 1.0;
      if (TS0) {
        c.mobs[jd] = mean[q * d + jd] + (-f);
      } else {
        const double rres = mean[q * d + jd] - f;
        double hm = mean[q * d + jd];
        for (int cc = 0; cc < q; ++cc) {
          for (int l = 0; l < d; ++l) {
            const double hv = -VF::jac(jd, cc, l, d, acc, params, tt);
            c.Hs[jd * hw + cc * d + l] = hv;
            hm = fma(hv, mean[cc * d + l], hm);
          }
        }
        c.mobs[jd] = hm + (rres - hm);
      }
    }
    __syncthreads();
  }

  // State columns <- the extrapolation stack [(A (pinv L))^T ; (s Q)^T] of the packed factor Lpk, A = kron(a, I_d),
  // Q = kron(q, diag(lam)); triangularise; scale column c by |p| so that rows 0..c hold column c of L_pred^T.
  PDEQ_DI static void extrapolate_chol(const Ctx& c, double (&col)[RPT], const double* Lpk, double s) {
    constexpr int d = D;
    if (c.is_state) {
      const int ci = c.ci, cj = c.cj, h = c.h;
      static_for<0, RPT>([&](auto ic_) {
        constexpr int i = decltype(ic_)::value;
        const int r = i * TPC + h;
        double val = 0.0;
        if (r < N) {
          // (A L~)[c][r] = sum_k a[ci][k] pinv_k L[k d + cj][r]   (k >= ci, and k d + cj >= r)
#pragma unroll
          for (int k = 0; k < n; ++k) {
            const int row = k * d + cj;
            if (k >= ci && row >= r) val = fma(c.A[ci][k], c.pinv[k] * Lpk[tri(row, r)], val);
          }
        } else if (r < 2 * N) {
          const int rr = r - N, ri = rr / d, rj = rr % d;
          val = (cj == rj && ci >= ri) ? s * c.Qm[ci][ri] * c.lam[cj] : 0.0;
        }
        col[i] = val;
      });
    }
    column_qr<2 * N, N, N, N>(col, c.pos_pred, c.h, [](int j) { return j * TPC; }, c.rbuf);
    if (c.is_state) {
      const double pc = fabs(c.p[c.ci]);
      static_for<0, RPT>([&](auto ic_) { col[decltype(ic_)::value] *= pc; });
    }
  }

  // With rows 0..N-1 of the state columns holding U = L^T (upper triangular), fill the observation columns with
  // [(H L)^T ; damp I] and triangularise the (N + d)-row stack, observation columns first. Afterwards R_Y is in RY,
  // R12 in rows 0..d-1 of the state columns, R_XY in their rows d..d+N-1.
  PDEQ_DI static void revert_stack(const Ctx& c, double (&col)[RPT], double damp) {
    constexpr int d = D, hw = HW;
    const int h = c.h, cp = c.cp, oa = c.oa;
    if (c.is_state && cp < hw) {  // publish rows 0..cp of the first hw columns of U
      static_for<0, (hw + TPC - 1) / TPC>([&](auto ic_) {
        constexpr int i = decltype(ic_)::value;
        const int r = i * TPC + h;
        if (r <= cp && r < hw) c.Ush[cp * hw + r] = col[i];
      });
    }
    if (c.is_state) {  // rows N.. of the state columns belong to the damp block: zero
      static_for<0, RPT>([&](auto ic_) {
        constexpr int i = decltype(ic_)::value;
        col[i] = (i * TPC + h >= N) ? 0.0 : col[i];
      });
    }
    __syncthreads();
    if (c.is_obs) {
      static_for<0, RPT>([&](auto ic_) {
        constexpr int i = decltype(ic_)::value;
        const int r = i * TPC + h;
        double val = 0.0;
        if (r < N) {
          // (H L)[a][r] = sum_e H[a][e] L[e][r], L[e][r] = U[r][e] (e >= r); H touches the first hw coefficients
          for (int e = r; e < hw; ++e) val = fma(c.Hs[oa * hw + e], c.Ush[e * hw + r], val);
        } else if (r == N + oa) {
          val = damp;
        }
        col[i] = val;
      });
    }
    column_qr<N + d, N + d, d, N>(col, c.pos_rev, c.h, [](int j) { return (j < D ? N + j : j - D) * TPC; }, c.rbuf);
    if (c.is_obs) {
      static_for<0, (d + TPC - 1) / TPC>([&](auto ic_) {
        constexpr int i = decltype(ic_)::value;
        const int r = i * TPC + h;
        if (r < d) c.RY[r * d + oa] = (r <= oa) ? col[i] : 0.0;
      });
    }
    __syncthreads();
  }

  // gain^T = R_Y^-1 R12 column by column (every lane of a state column solves its own d x d system), then
  // mean_out = mean_in - gain mobs. `lstsq`: a zero pivot gives a zero gain component (the minimum-norm solution of
  // linalg.lstsq_svd when the zero pivots are decoupled) -- only the update at t0 can meet one.
  PDEQ_DI static void apply_gain(const Ctx& c, const double (&col)[RPT], const double* mean_in, double* mean_out,
                                 bool lstsq) {
    constexpr int d = D;
    if (c.is_state) {
      double r12[d];
      const int base = (c.tid & 31) & ~(TPC - 1);
      static_for<0, (d + TPC - 1) / TPC>([&](auto ic_) {
        constexpr int i = decltype(ic_)::value;
        static_for<0, TPC>([&](auto pc_) {
          constexpr int part = decltype(pc_)::value;
          const double x = __shfl_sync(c.gmask, col[i], base + part);
          if constexpr (i * TPC + part < d) r12[i * TPC + part] = x;
        });
      });
      static_for<0, d>([&](auto ic_) {
        constexpr int i = d - 1 - decltype(ic_)::value;
        double acc = r12[i];
        static_for<i + 1, d>([&](auto lc_) {
          constexpr int l = decltype(lc_)::value;
          acc = fma(-c.RY[i * d + l], r12[l], acc);
        });
        const double piv = c.RY[i * d + i];
        r12[i] = (lstsq && piv == 0.0) ? 0.0 : acc * fast_rcp(piv);
      });
      double corr = 0.0;
      static_for<0, d>([&](auto ac_) {
        constexpr int a_ = decltype(ac_)::value;
        corr = fma(r12[a_], c.mobs[a_], corr);
      });
      if (c.h == 0) mean_out[c.cp] = mean_in[c.cp] - corr;
    }
    __syncthreads();
  }

  // packed factor <- rows row0..row0+c of state column c (transposed): the accepted R_XY^T, or L_pred^T
  PDEQ_DI static void store_factor(const Ctx& c, const double (&col)[RPT], double* Lpk, int row0) {
    if (c.is_state) {
      const int h = c.h, cp = c.cp;
      static_for<0, RPT>([&](auto ic_) {
        constexpr int i = decltype(ic_)::value;
        const int jj = i * TPC + h - row0;
        if (jj >= 0 && jj <= cp) Lpk[tri(cp, jj)] = col[i];
      });
    }
  }

  // whitened residual of an observation with upper-triangular factor RY (R^T is the left square root): solve
  // R^T w = mobs, rms; optionally the row norms of R^T (the marginal standard deviations). One thread.
  PDEQ_DI static double whiten(const Ctx& c, bool want_std) {
    constexpr int d = D;
    double ss = 0.0;
    for (int i = 0; i < d; ++i) {
      double acc = c.mobs[i];
      for (int l = 0; l < i; ++l) acc = fma(-c.RY[l * d + i], c.wht[l], acc);
      c.wht[i] = acc * fast_rcp(c.RY[i * d + i]);
      ss = fma(c.wht[i], c.wht[i], ss);
      if (want_std) {
        double rn = 0.0;
        for (int l = 0; l <= i; ++l) rn = fma(c.RY[l * d + i], c.RY[l * d + i], rn);
        c.stdv[i] = safe_sqrt(rn);
      }
    }
    return safe_sqrt(ss) * c.inv_sqrt_d;
  }

  PDEQ_DI static void run(const LoopArgs& a, double* __restrict__ smem) {
    const pdeq_config& cfg = a.cfg;
    const double(*__restrict__ A)[PDEQ_MAX_COEFFS] = cfg.sys_a;
    const double(*__restrict__ Qm)[PDEQ_MAX_COEFFS] = cfg.sys_q;
    const double* __restrict__ fact = cfg.factorials;
    const double* __restrict__ ifact = cfg.inv_factorials;
    const bool adaptive = a.fixed_grid == 0;
    const bool clip = cfg.clip_dt != 0;
    const bool needs_interp = adaptive && !clip;
    const int T = a.T;
    constexpr int d = D, hw = HW;
    const long B = a.prob.num_instances;
    const int max_attempts = cfg.max_attempts > 0 ? cfg.max_attempts : 0x7fffffff;
    const double inv_sqrt_d = rsqrt((double)d);
    const double neg_inv_n = -1.0 / (double)n;
    const int tid = threadIdx.x;
    const DenseSmemLayout lay = DenseSmemLayout::make(n, d, q, needs_interp);

    // this lane's column: pairs 0..N-1 hold the state columns, pairs N..N+d-1 the observation columns
    const int cp = tid / TPC, h = tid % TPC;
    const bool is_state = cp < N, is_obs = cp >= N && cp < LD;
    const int ci = is_state ? cp / d : 0, cj = is_state ? cp % d : 0;  // coefficient / dimension of a state column
    const int oa = is_obs ? cp - N : 0;                                 // observation index of an observation column
    const unsigned gmask = ((1u << TPC) - 1u) << ((tid & 31) & ~(TPC - 1));
    const int pos_pred = is_state ? cp : -1;                    // extrapolation stack: state columns only
    const int pos_obs = is_obs ? oa : -1;                       // observation-only stack
    const int pos_rev = is_state ? d + cp : (is_obs ? oa : -1);  // correction stack: observation columns first

    double* Lfrom = smem + lay.off_Lfrom;
    double* Lif = smem + lay.off_Lif;
    double* v = smem + lay.off_vec;
    double* rbuf = v;                  // 2 x RB
    double* m_from = rbuf + 2 * RB;
    double* mp = m_from + N;
    double* m_new = mp + N;
    double* m_if = m_new + N;
    double* Hs = m_if + N;             // d x (q+1) d: the non-zero columns of the linearisation
    double* Ush = Hs + d * hw;         // hw x hw: rows 0..e of the first hw columns of the factor, for (H L)^T
    double* RY = Ush + hw * hw;        // d x d: R_Y (correction) / R_obs (observation-only), row-major upper
    double* mobs = RY + d * d;         // d
    double* wht = mobs + d;            // d: whitened residual
    double* stdv = wht + d;            // d: error estimate per dimension
    double* refv = stdv + d;           // d
    double* lam = refv + d;            // d: prior output scale (diagonal of Lambda)
    double* p = lam + d;               // 8
    double* pinv = p + 8;              // 8
    double* red = pinv + 8;            // 8
    double* bc = red + 8;              // 8: broadcast slots

    double col[RPT];  // this lane's rows of its column
    double params[P];
    double t = 0.0, dt = 0.0, ctrl_lprev = 0.0, ndata = 0.0, t_next = 0.0, t_if = 0.0, sig = 1.0, run_scale = 0.0;
    int nsteps = 0, nattempts = 0, ck = 0, status = 0;
    long b = -1;
    bool need_load = true;

    Ctx cx;
    cx.tid = tid; cx.cp = cp; cx.h = h; cx.ci = ci; cx.cj = cj; cx.oa = oa;
    cx.pos_pred = pos_pred; cx.pos_obs = pos_obs; cx.pos_rev = pos_rev;
    cx.is_state = is_state; cx.is_obs = is_obs; cx.gmask = gmask; cx.inv_sqrt_d = inv_sqrt_d;
    cx.A = A; cx.Qm = Qm; cx.rbuf = rbuf; cx.Hs = Hs; cx.Ush = Ush; cx.RY = RY; cx.mobs = mobs; cx.wht = wht;
    cx.stdv = stdv; cx.lam = lam; cx.p = p; cx.pinv = pinv;

    static_for<0, RPT>([&](auto ic_) { col[decltype(ic_)::value] = 0.0; });

    while (true) {
      if (need_load) {
        __syncthreads();
        if (tid == 0) reinterpret_cast<long*>(bc)[0] = (long)atomicAdd(a.work_counter, 1ULL);
        __syncthreads();
        b = reinterpret_cast<long*>(bc)[0];
        if (b >= B) break;
        if (a.prob.order != nullptr) b = (long)a.prob.order[b];  // service order (pdeq_problem.order)
        need_load = false;
#pragma unroll
        for (int k = 0; k < P; ++k)
          params[k] = (VF::num_params > 0) ? a.prob.params[b * a.prob.params_stride + k] : 0.0;
        for (int e = tid; e < N; e += G) m_from[e] = a.prob.tcoeffs[b * N + e];
        for (int e = tid; e < TRI_N; e += G) Lfrom[e] = 0.0;
        __syncthreads();
        if (a.prob.init_std != nullptr) {
          const double* sd = a.prob.init_std + b * a.prob.init_std_stride;
          for (int e = tid; e < N; e += G) Lfrom[tri(e, e)] = sd[e];
        }
        for (int e = tid; e < d; e += G)
          lam[e] = a.prob.prior_scale != nullptr ? a.prob.prior_scale[b * a.prob.prior_scale_stride + e] : 1.0;
        t = a.grid[0];
        dt = adaptive ? a.dt0[b * a.dt0_stride] : 0.0;
        ctrl_lprev = 0.0;
        ndata = 0.0;
        nsteps = 0;
        nattempts = 0;
        status = 0;
        sig = 1.0;
        run_scale = 0.0;
        __syncthreads();
        if (cfg.constraint_init != 0) {
          // solver.init with constraint_init (solvers.py:361-372, 526-537, 670-680): condition the initial state on a
          // zero residual of the constraint linearised at it
          if (is_state) {
            static_for<0, RPT>([&](auto ic_) {
              constexpr int i = decltype(ic_)::value;
              const int r = i * TPC + h;  // U = L^T of the (diagonal) initial factor
              col[i] = (r <= cp && r < N) ? Lfrom[tri(cp, r)] : 0.0;
            });
          }
          linearise_at(cx, m_from, t, params);
          revert_stack(cx, col, a.damp);
          apply_gain(cx, col, m_from, m_new, true);
          if (cfg.solver == PDEQ_SOLVER_MLE) {
            // solver_mle.init (solvers.py:361-374): the update at t0 is the first datum of the running calibration
            if (tid == 0) bc[2] = whiten(cx, false);
            __syncthreads();
            run_scale = bc[2];
            ndata = 1.0;
          }
          for (int e = tid; e < N; e += G) m_from[e] = m_new[e];
          store_factor(cx, col, Lfrom, d);
          __syncthreads();
        }
        emit(a, b, 0, t, m_from, Lfrom, 1.0, 1.0, 0);
        if (needs_interp) {
          for (int e = tid; e < N; e += G) m_if[e] = m_from[e];
          for (int e = tid; e < TRI_N; e += G) Lif[e] = Lfrom[e];
        }
        t_if = t;
        ck = 1;
        t_next = (T > 1) ? a.grid[1] : t;
        __syncthreads();
      }

      const bool at_checkpoint = (ck >= T) || (adaptive && !(t + a.eps < t_next));
      if (at_checkpoint) {
        if (ck < T) {
          if (needs_interp && t > t_next + a.eps) {
            // interp_beyond_t1: extrapolate interp_from to t_next with the scale of the overstepped state
            const double dti = t_next - t_if;
            if (tid == 0) {
              double pp[n], pi[n];
              preconditioner<n>(dti, ifact, fact, pp, pi);
#pragma unroll
              for (int i = 0; i < n; ++i) {
                p[i] = pp[i];
                pinv[i] = pi[i];
              }
            }
            __syncthreads();
            for (int e = tid; e < N; e += G) {
              const int i = e / d, jd = e % d;
              double acc = 0.0;
              for (int k = i; k < n; ++k) acc = fma(A[i][k], pinv[k] * m_if[k * d + jd], acc);
              mp[e] = p[i] * acc;
            }
            extrapolate_chol(cx, col, Lif, safe_sqrt(fabs(dti)) * sig);
            __syncthreads();
            store_factor(cx, col, Lif, 0);
            for (int e = tid; e < N; e += G) m_if[e] = mp[e];
            __syncthreads();
            emit(a, b, ck, t_next, m_if, Lif, 1.0, sig, nsteps);
            t_if = t_next;
          } else {
            emit(a, b, ck, t, m_from, Lfrom, 1.0, sig, nsteps);
            if (needs_interp) {
              for (int e = tid; e < N; e += G) m_if[e] = m_from[e];
              for (int e = tid; e < TRI_N; e += G) Lif[e] = Lfrom[e];
            }
            t_if = t;
          }
          ck += 1;
          if (ck < T) t_next = a.grid[ck];
        }
        if (ck >= T) {
          if (cfg.solver == PDEQ_SOLVER_MLE) {
            double fin = run_scale;
            if (cfg.correct_asymptotic_underconfidence) fin = fin / sqrt((double)nsteps);
            __syncthreads();
            for (int c = 0; c < T; ++c) {
              const long bt = b * T + c;
              if (a.sol.chol != nullptr) {
                double* co = a.sol.chol + bt * (long)N * N;
                for (int e = tid; e < N * N; e += G) co[e] = fin * co[e];
              }
              if (tid == 0 && a.sol.output_scale != nullptr) a.sol.output_scale[bt] = fin;
            }
          }
          double bad = 0.0;
          for (int e = tid; e < N; e += G) bad += isfinite(m_from[e]) ? 0.0 : 1.0;
          bad = block_sum(bad, red);
          if (status == 0 && bad > 0.0) status = PDEQ_STATUS_NONFINITE;
          if (tid == 0) {
            a.sol.status[b] = status;
            if (a.sol.num_attempts != nullptr) a.sol.num_attempts[b] = nattempts;
          }
          need_load = true;
        }
        __syncthreads();
        continue;
      }

      // ------------------------------------------------------------------ one step attempt
      nattempts += 1;
      double dtc;
      if (adaptive) {
        dtc = clip ? fmin(dt, t_next - t) : dt;
      } else {
        dtc = a.grid[ck] - a.grid[ck - 1];
      }
      const double sq = safe_sqrt(fabs(dtc));
      const double t_new = t + dtc;
      __syncthreads();
      if (tid == 0) {
        double pp[n], pi[n];
        preconditioner<n>(dtc, ifact, fact, pp, pi);
#pragma unroll
        for (int i = 0; i < n; ++i) {
          p[i] = pp[i];
          pinv[i] = pi[i];
        }
      }
      __syncthreads();
      for (int e = tid; e < N; e += G) {
        const int i = e / d, jd = e % d;
        double acc = 0.0;
        for (int k = i; k < n; ++k) acc = fma(A[i][k], pinv[k] * m_from[k * d + jd], acc);
        mp[e] = p[i] * acc;
      }
      __syncthreads();

      linearise_at(cx, mp, t_new, params);

      // observation of the zero-error extrapolation: R_obs = qr_r([(H L_u)^T ; damp I]) (solver_dynamic, residual error)
      const bool need_obs = adaptive ? (cfg.solver == PDEQ_SOLVER_DYNAMIC || cfg.error == PDEQ_ERROR_RESIDUAL_STD)
                                     : (cfg.solver == PDEQ_SOLVER_DYNAMIC);
      double sig_new = 1.0, whitened_obs = 0.0;
      if (need_obs) {
        if (is_obs) {
          static_for<0, RPT>([&](auto ic_) {
            constexpr int i = decltype(ic_)::value;
            const int r = i * TPC + h;
            double val = 0.0;
            if (r < N) {
              // (H L_u)[a][k d + l] = sum_{i >= k} H[a][i d + l] |p_i| sq q[i][k] lam_l
              const int k = r / d, l = r % d;
              for (int ii = k; ii <= q && ii < n; ++ii) val = fma(Hs[oa * hw + ii * d + l], fabs(p[ii]) * sq * Qm[ii][k], val);
              val *= lam[l];
            } else if (r == N + oa) {
              val = a.damp;
            }
            col[i] = val;
          });
        }
        column_qr<N + d, d, d, N>(col, pos_obs, h, [](int j) { return (N + j) * TPC; }, rbuf);
        if (is_obs) {
          static_for<0, (d + TPC - 1) / TPC>([&](auto ic_) {
            constexpr int i = decltype(ic_)::value;
            const int r = i * TPC + h;
            if (r < d) RY[r * d + oa] = (r <= oa) ? col[i] : 0.0;
          });
        }
        __syncthreads();
        if (tid == 0) bc[1] = whiten(cx, true);
        __syncthreads();
        whitened_obs = bc[1];
        if (cfg.solver == PDEQ_SOLVER_DYNAMIC) sig_new = whitened_obs;
      }

      if (adaptive && cfg.error != PDEQ_ERROR_RESIDUAL_STD) {
        // error_state_std (solvers.py:1070-1086): Bayes rule on the zero-error extrapolation. Done first because it
        // needs the column registers, which afterwards hold the proposal until the step is accepted or rejected.
        const int idx = cfg.derivative_idx;
        if (is_state) {
          static_for<0, RPT>([&](auto ic_) {
            constexpr int i = decltype(ic_)::value;
            const int r = i * TPC + h;  // U = L_u^T: U[r][c] = L_u[c][r], block lower in (ci >= ri), same dimension
            const int ri = r / d, rj = r % d;
            col[i] = (r < N && cj == rj && ci >= ri) ? fabs(p[ci]) * sq * Qm[ci][ri] * lam[cj] : 0.0;
          });
        }
        revert_stack(cx, col, a.damp);
        if (tid == 0) bc[3] = whiten(cx, false);
        __syncthreads();
        if (is_state && ci == idx) {
          // std of coefficient idx, dimension cj: norm of rows d..d+cp of this column (R_XY)
          double rn = 0.0;
          static_for<0, RPT>([&](auto ic_) {
            constexpr int i = decltype(ic_)::value;
            const int r = i * TPC + h;
            rn = (r >= d && r <= d + cp) ? fma(col[i], col[i], rn) : rn;
          });
          rn = group_sum(rn, gmask);
          if (h == 0) stdv[cj] = bc[3] * safe_sqrt(rn);
        }
        __syncthreads();
      }

      // extrapolate the factor, correct; the proposal's factor stays in the column registers until the step is
      // accepted or rejected
      extrapolate_chol(cx, col, Lfrom, sq * sig_new);
      revert_stack(cx, col, a.damp);
      apply_gain(cx, col, mp, m_new, false);
      double run_new = run_scale;
      if (cfg.solver == PDEQ_SOLVER_MLE) {
        if (tid == 0) bc[2] = whiten(cx, false);
        __syncthreads();
        const double w1 = sqrt(ndata / (ndata + 1.0)), w2 = sqrt(1.0 / (ndata + 1.0));
        const double x1 = w1 * run_scale, x2 = w2 * bc[2];
        run_new = safe_sqrt(fma(x1, x1, x2 * x2));
      }
      __syncthreads();

      // ------------------------------------------------------------------ error estimate + control
      bool accept = true;
      double dt_next = dt;
      if (adaptive) {
        int kpow, idx = 0;
        if (cfg.error == PDEQ_ERROR_RESIDUAL_STD) {
          // error_d = sigma * std_d of the observed marginal (solvers.py:955-960); stdv holds the row norms
          for (int e = tid; e < d; e += G) {
            stdv[e] = whitened_obs * stdv[e];
            refv[e] = fmax(fabs(m_from[e]), fabs(m_new[e]));
          }
          kpow = q;
        } else {
          idx = cfg.derivative_idx;  // stdv was computed ahead of the extrapolation
          for (int e = tid; e < d; e += G) refv[e] = fmax(fabs(m_from[idx * d + e]), fabs(m_new[idx * d + e]));
          kpow = idx;
        }
        if (cfg.error_per_unit_step) kpow += 1;
        __syncthreads();
        if (tid == 0) {
          double escale = ipow_small<n>(dtc, kpow);
          for (int e = 0; e <= n; ++e) {
            if (e == kpow) escale *= ifact[e];
          }
          double norm;
          if (cfg.error_norm == PDEQ_NORM_SCALE_THEN_RMS) {
            double ss = 0.0;
            for (int e = 0; e < d; ++e) {
              const double w = (stdv[e] * escale) * fast_rcp(fma(a.rtol, refv[e], a.atol));
              ss = fma(w, w, ss);
            }
            norm = safe_sqrt(ss) * inv_sqrt_d;
          } else {
            double se2 = 0.0, sr2 = 0.0;
            for (int e = 0; e < d; ++e) {
              const double ea = stdv[e] * escale;
              se2 = fma(ea, ea, se2);
              sr2 = fma(refv[e], refv[e], sr2);
            }
            norm = (safe_sqrt(se2) * inv_sqrt_d) * fast_rcp(fma(a.rtol, safe_sqrt(sr2) * inv_sqrt_d, a.atol));
          }
          bc[4] = neg_inv_n * log2(norm);
        }
        __syncthreads();
        const double lep = bc[4];
        accept = !(lep < 0.0);
        double lratio;
        if (cfg.control == PDEQ_CONTROL_PI) {
          lratio = fma(cfg.exponent_integral, lep, cfg.exponent_proportional * (lep - ctrl_lprev));
          if (lep >= 0.0) ctrl_lprev = lep;
        } else {
          lratio = lep;
        }
        const double ratio = cfg.safety * exp2(lratio);
        dt_next = fmax(cfg.factor_min, fmin(ratio, cfg.factor_max)) * dtc;
        if (a.sol.trace != nullptr && nattempts <= a.sol.trace_capacity && tid == 0) {
          double* tr = a.sol.trace + (b * a.sol.trace_capacity + (nattempts - 1)) * 4;
          tr[0] = t;
          tr[1] = dtc;
          tr[2] = exp2(lep);
          tr[3] = accept ? 1.0 : 0.0;
        }
        if (nattempts >= max_attempts) {
          status = PDEQ_STATUS_MAX_ATTEMPTS;
          accept = true;
          ck = T;
        }
      }

      // ------------------------------------------------------------------ commit
      dt = dt_next;
      if (accept) {
        if (needs_interp) {
          for (int e = tid; e < N; e += G) m_if[e] = m_from[e];
          for (int e = tid; e < TRI_N; e += G) Lif[e] = Lfrom[e];
          t_if = t;
        }
        __syncthreads();
        for (int e = tid; e < N; e += G) m_from[e] = m_new[e];
        store_factor(cx, col, Lfrom, d);
        if (cfg.solver == PDEQ_SOLVER_DYNAMIC) sig = sig_new;
        run_scale = run_new;
        ndata += 1.0;
        t = t_new;
        nsteps += 1;
        __syncthreads();
        if (!adaptive) {
          emit(a, b, ck, t, m_from, Lfrom, 1.0, sig, nsteps);
          ck += 1;
        }
      }
    }
  }
};

template <class VF, int NU, bool TS0>
__global__ void __launch_bounds__(DenseSmemLayout::threads((NU + 1) * VF::fixed_dim, VF::fixed_dim), PDEQ_K3_MIN_BLOCKS)
    k3_loop_kernel(const __grid_constant__ LoopArgs a) {
  extern __shared__ double smem_k3[];
  DenseLoop<VF, NU, TS0>::run(a, smem_k3);
}

}  // namespace pdeq
