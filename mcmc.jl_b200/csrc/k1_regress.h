#pragma once
#include "common.cuh"
namespace mg {

// K1: fused likelihood + gradient for the regression families (linear / logistic / probit).
constexpr int K1_ROWS = 32;      // rows of X per shared-memory tile
constexpr int K1_WARPS = 8;      // consumer warps per CTA; each owns 8 chains
constexpr int K1_CHAINS = 8 * K1_WARPS;  // chains per CTA (64)
constexpr int K1_MAX_DK_3CTA = 4;   // d <= 32: small accumulators, three CTAs per SM hide the epilogue latency
constexpr int K1_MAX_DK_2CTA = 13;  // d <= 104: 128 registers per thread, two CTAs per SM
constexpr int K1_MAX_D = 200;

// Packed design matrix: tile t holds rows [32t, 32t+32) as a ready-made shared-memory image:
//   double x[32][S]  (row-major, S = 8*DK + 4 so that S = 4 (mod 8): conflict-free fragment loads)
//   double y[32]
// One cp.async.bulk per tile brings it in.  Rows >= N are zero.
struct K1Pack {
  double* tiles = nullptr;
  int64_t N = 0, d = 0;
  int DK = 0;        // ceil(d/8)
  int S = 0;         // row stride in doubles
  int64_t ntiles = 0;
  int64_t tile_doubles = 0;
  const unsigned int* ynb = nullptr;   // device word behind the tiles: nonzero iff some response is neither 0.0 nor 1.0
};

struct K1Args {
  K1Pack P;
  int32_t family;
  double hyper[4];
  const double* q;       // [d][Cp] evaluation points
  double* part;          // [nsplit][d+2][Cp]: rows 0..d-1 X'r, row d loglik, row d+1 non-finite count
  int64_t Cp;
  int32_t nsplit;
  int32_t need_grad;     // 0: log-likelihood only (RWM)
  const uint8_t* need_ll;  // [Cp] per-chain flag, or null = all chains need the log-likelihood
  const int32_t* phase;    // [Cp] per-chain phase (PH_DONE chains are skipped), or null
  const int32_t* remaining; // device counter of unfinished chains, or null
  int32_t debug;           // experiments only: 1 = skip the link function (r = eta), 2 = skip phase 2
  // fused interior leapfrog (nsplit == 1 only: the CTA then holds the complete gradient of its 64 chains in registers).
  // A chain in the middle of an HMC / HMCDA trajectory (phase PH_LEAP, leap + 2 <= nleaps_cur) gets the end of this
  // leapfrog and the start of the next one (HMC.jl:98,95,96) right here instead of a round trip of X'r through `part`
  // and a pass of the transition kernel over mom / q; the transition kernel then only advances the chain's counters.
  int32_t fuse_leap;
  int32_t* leap;             // [Cp]  advanced here for the chains whose leapfrog this kernel completes
  uint8_t* need_ll_rw;       // [Cp]  their flag for the NEXT evaluation (the trajectory's last one needs the value)
  uint8_t* k1_done;          // [Cp]  set for those chains: the transition kernel has nothing left to do for them this wave
  unsigned long long* n_evals;   // evaluation counter of the run (one per such chain is added here)
  const int32_t* nleaps_cur; // [Cp]
  const double* eps_cur;     // [Cp]
  double* mom;               // [d][Cp]
  double* q_rw;              // == q
};

cudaError_t k1_pack(K1Pack& P, const double* dX /* N x d col-major, device */, const double* dy, int64_t N, int64_t d,
                    cudaStream_t st);
void k1_free(K1Pack& P);
int k1_choose_splits(const K1Pack& P, int64_t Cp, int device = -1, int family = -1);   // device < 0: the current device; the resident CTA count of the small-d class depends on the family
cudaError_t k1_launch(const K1Args& a, cudaStream_t st);
bool k1_supported(int64_t d);

}  // namespace mg
