/* libsurreal_b200 -- C-ABI of the B200-native actor -> replay -> learner hot path.
 *
 * The reference (SurrealAI/surreal) is pure Python and has NO FFI / operator registry: its plugin
 * surface is three Python base classes (Agent / Replay / Learner).  This header is
 * therefore the boundary a maintainer would bind from those classes (ctypes stub in
 * INTEGRATION.md); every entry point cites the reference code whose arithmetic it replaces.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _h; fp32, row-major, contiguous
 *     unless a leading dimension is given; `stream` is a cudaStream_t passed as void*.
 *   - every call is asynchronous on `stream`, allocates nothing, and returns SB200_OK (0) or a
 *     negative sb200 status; no exceptions cross the ABI.  Workspaces are caller-owned.
 *   - weights use the KERNEL LAYOUT: W[l] is [in_l][ldw_l] (ldw = out rounded up to 4), i.e. the
 *     transpose of torch.nn.Linear.weight, zero in the padding columns; b[l] is [ldw_l].
 *   - there is no CPU fallback: without a CUDA device every compute entry fails.
 */
#ifndef SURREAL_B200_H
#define SURREAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB200_OK 0
#define SB200_ERR_ARG (-1)
#define SB200_ERR_CUDA (-2)
#define SB200_ERR_UNSUPPORTED (-3)

#define SB200_ACT_NONE 0
#define SB200_ACT_RELU 1
#define SB200_ACT_TANH 2
#define SB200_MAX_LAYERS 4

int sb200_version(void);
const char* sb200_status_string(int status);
/* SM count / compute capability of the current device (fails without a GPU). */
int sb200_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------
 * Networks.  One descriptor covers PPO_ActorNetwork / PPO_CriticNetwork
 * (surreal/model/model_builders/builders.py:86-175), ActorNetworkX / CriticNetworkX (:35-84).
 * `aux_layer` >= 0 concatenates `aux` columns to the input of that layer (the DDPG critic's
 * cat(h, action), builders.py:80-83). */
typedef struct {
    int n_layers;                       /* 1..SB200_MAX_LAYERS */
    int dims[SB200_MAX_LAYERS + 1];     /* dims[0] = input width (without aux), dims[l+1] = out of layer l */
    int act[SB200_MAX_LAYERS];          /* SB200_ACT_* applied after layer l */
    const float* W[SB200_MAX_LAYERS];   /* [in_l(+aux)][ldw_l] */
    const float* b[SB200_MAX_LAYERS];   /* [ldw_l] */
    int ldw[SB200_MAX_LAYERS];
    int aux_layer;                      /* -1: none */
    int aux_dim;
} sb200_mlp;

/* ZFilter (surreal/model/z_filter.py:59-79): stats = running_sum[D] | running_sumsq[D] | count[1]. */
typedef struct {
    const float* stats;                 /* NULL: no filter */
    float eps;                          /* 1e-5 in the reference */
} sb200_zfilter;

/* Input rows of a forward pass.  Either a plain matrix (x, ldx, rows) or, when win_n > 0, the
 * virtual concatenation cat([obs, obs_next], dim=1).view(-1, D) of ppo.py:376-383 WITHOUT copying:
 * row r -> (b, k) = (r / (win_n+1), r % (win_n+1)); k < win_n ? x[(b*win_n+k)*ldx] : x_next[b*ldx]. */
typedef struct {
    const float* x;
    const float* x_next;                /* only when win_n > 0 */
    int64_t ldx;
    int64_t rows;                       /* total rows ( = B*(win_n+1) when win_n > 0 ) */
    int win_n;
    const float* aux;                   /* [rows][aux_ld] or NULL */
    int64_t aux_ld;
} sb200_rows;

/* Fused forward of a whole MLP on row tiles (z-filter -> Linear/act x n_layers): activations stay in
 * shared memory; `save[l]` (may be NULL) receives layer l's post-activation output [rows][ld_save[l]]
 * (needed by the backward pass); save[n_layers-1] is the network output.
 * Replaces: ppo_net.py:253-315 (forward_actor / forward_critic), builders.py:114-132,160-175,
 * ddpg_net.py:63-91. */
int sb200_mlp_forward_f32(const sb200_mlp* net, const sb200_zfilter* zf, const sb200_rows* in,
                          float* const* save, const int64_t* ld_save, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Windowed GAE + n-step return (surreal/learner/ppo.py:372-374,387-418).
 *   rewards [B,n], values [B,n+1] (raw critic output; the (1-done) mask of ppo.py:387 is applied
 *   inside), dones [B,n].  horizon == n -> MLP branch (one output per window), horizon < n ->
 *   RNN branch (E = n-horizon+1 outputs per window).  adv/ret are [B,E].  norm_adv applies
 *   (adv-mean)/max(unbiased_std, 1e-4) over all B*E advantages.
 *   workspace: sb200_gae_workspace_bytes() bytes, zero-initialised once by the caller. */
size_t sb200_gae_workspace_bytes(int B, int n, int horizon);
int sb200_gae_window_f32(const float* rewards, const float* values, const float* dones, int B, int n,
                         int horizon, double gamma, double lam, int norm_adv, float* adv, float* ret,
                         void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SURREAL_B200_H */
