/* pvcr_b200 — C ABI of the B200-native captioning hot path.
 *
 * Drop-in boundary for p-kar/pytorch-video-caption-rationale.  The reference has no FFI: its boundary is
 * the Python nn.Module surface (model/S2VTAttModel.py:199-264, model/S2VTModel.py:12-202,
 * model/RationaleNet.py:57-106) plus the loss contract (train_utils.py:22-95).  Each entry point below
 * names the reference call it replaces; INTEGRATION.md shows the ctypes binding.
 *
 * Conventions: every pointer is a DEVICE pointer unless stated; matrices are row-major fp32 with an
 * explicit leading dimension in elements; token ids / lengths are int64 (torch.long); `stream` is a
 * cudaStream_t passed as void*.  No hidden allocation: scratch comes from the caller through
 * (workspace, workspace_bytes) sized by the matching *_workspace query.  All work is enqueued on
 * `stream`; nothing synchronises.  Return value 0 = ok, negative = error (pvcr_last_error()).
 *
 * `nsplit` selects the tensor-core arithmetic: 1 = plain bf16 operands with fp32 accumulation (training
 * throughput mode), 2 / 3 = each fp32 operand split into 2 / 3 bf16 terms (3 / 6 tcgen05 products per
 * logical product; nsplit = 3 reproduces fp32 arithmetic and is what greedy decoding uses so that token ids
 * match the fp32 reference bit for bit).
 */
#ifndef PVCR_B200_H
#define PVCR_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* pvcr_last_error(void);
int pvcr_version(void);

/* Launch accounting used by bench.py: every kernel launch of the library is counted per kernel class; with
 * pvcr_prof_enable(1) each launch is additionally bracketed by a CUDA-event pair on its stream (2: also launches
 * captured into a CUDA graph, as external event-record nodes that every replay re-records).
 * pvcr_prof_read fills launches[], ms[] (summed event durations; synchronises) and work[] (executed tensor-core
 * FLOPs for the GEMM class) for pvcr_prof_num_classes() classes, all since the last pvcr_prof_reset(). */
int pvcr_prof_num_classes(void);
const char* pvcr_prof_class_name(int cls);
void pvcr_prof_enable(int on);
void pvcr_prof_reset(void);
int pvcr_prof_read(uint64_t* launches, double* ms, double* work);
/* Timeline of the event-timed launches since the last reset (host enqueue order): class, start and end in ms relative
 * to the first launch; returns the number of entries written (<= cap) or a negative error code. */
int pvcr_prof_timeline(int* cls, float* t0_ms, float* t1_ms, int cap);
/* Every event-timed launch since the last reset (host enqueue order): class, duration, work (executed tensor-core FLOPs
 * of a GEMM launch, else 0); returns the number of entries written (<= cap) or a negative error code. */
int pvcr_prof_launch_list(int* cls, float* ms, double* work, int cap);

/* Registers a device-resident uint64 counter that every dropout / Gumbel draw mixes into its seed when the kernel
 * runs (NULL switches it off).  CUDA-graph replays otherwise repeat the seed baked in at capture; with the counter
 * incremented once per step (pvcr_b200.graphs does it inside the captured graph) every replay draws fresh masks,
 * forward and backward of one step still agreeing. */
void pvcr_set_seed_step(const uint64_t* device_counter);

/* Side lane: the library owns a second, lower-priority stream for work that is off the step's critical path (weight
 * gradients, bias column sums, the embedding scatter), so that it runs next to the persistent recurrent sweeps, which
 * leave 20 of the 148 SMs free, instead of between them.  Fork and join are event edges (CUDA-graph capturable).
 *   mode 0: off - everything on the caller's stream;
 *   mode 1: (default) fork and join inside each call - outputs are complete in stream order when a call returns;
 *   mode 2: joins deferred - outputs and workspaces of the *_bwd calls are only safe after pvcr_side_join(stream),
 *           which the caller must issue before consuming gradients, ending a stream capture or freeing workspaces.
 * pvcr_side_mode returns the previous mode (a value outside 0..2 only queries). */
int pvcr_side_mode(int mode);
int pvcr_side_join(void* stream);
/* `stream` waits for ONE lane's current point only (lane in 0..2); the other lanes stay un-joined.  Which lane produces
 * which gradients in mode 2: after pvcr_s2vtatt_bwd_part(part = 1) lane 1 carries d W_ih(dec) / d embedding / d b_ih(dec),
 * lane 0 d W_hh(dec) / d W_q, lane 2 d v / d b_hh(dec) / d W_k. */
int pvcr_side_join_lane(void* stream, int lane);
/* `stream` waits for a milestone recorded in the middle of a lane's work.  id 0: the embedding gradient written by the
 * latest pvcr_s2vtatt_bwd_part(part = 1) is final (its lane goes on with other weight gradients) -- the point at which a
 * data-parallel caller can start all-reducing the largest decoder-side gradient. */
int pvcr_side_wait_milestone(void* stream, int id);

/* Tuning aid: in-kernel phase timestamps (clock64 of CTA 0, [step][8]) of the last persistent-kernel launch. */
int pvcr_debug_phase_timing(int on);
int pvcr_debug_phase_read(long long* out, int steps);

/* y[M,N] = x[M,K] w[N,K]^T + bias[N]      (torch.nn.Linear / F.linear; bias may be NULL) */
size_t pvcr_linear_fwd_workspace(int M, int N, int K, int nsplit);
int pvcr_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, float* y,
                    int64_t ldy, int M, int N, int K, int nsplit, void* workspace, size_t workspace_bytes,
                    void* stream);
/* dx = dy w ; dw (+)= dy^T x ; db (+)= colsum(dy)   (autograd of F.linear; any output may be NULL) */
size_t pvcr_linear_bwd_workspace(int M, int N, int K, int nsplit);
int pvcr_linear_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* w, int64_t ldw,
                    float* dx, int64_t lddx, float* dw, int64_t lddw, float* db, int M, int N, int K, int nsplit,
                    int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* dw[N,K] (+)= dy[R,N]^T x[R,K] in bf16 arithmetic with MN-major tcgen05 operands: the operands are cast to bf16
 * row-major as they are (no transposed copies), the contraction runs over their rows (autograd of F.linear w.r.t.
 * the weight; the building block of every hoisted weight gradient). */
size_t pvcr_wgrad_mn_workspace(int R, int N, int K);
int pvcr_wgrad_mn(const float* dy, int64_t lddy, const float* x, int64_t ldx, float* dw, int64_t lddw, int R, int N,
                  int K, int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Model-level entry points.  Symbols: B batch, N frames, V feature size, H hidden, E embedding size,
 * L = max_len, Vc vocabulary.  dropout_p / seed parameterise the counter-based (Philox) dropout masks of
 * nn.Dropout call sites; parity tests run with dropout_p = 0 as the reference draws from torch's RNG.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int B, N, V, H, E, L, Vc;
  int nsplit;
  float dropout_p;
  uint64_t seed;
} PvcrDims;

/* S2VTAttModel parameters (reference state_dict names, model/S2VTAttModel.py:60-61,100-123). */
typedef struct {
  const float* enc_w_ih; /* encoder.rnn.weight_ih_l0            [3H, V]   */
  const float* enc_w_hh; /* encoder.rnn.weight_hh_l0            [3H, H]   */
  const float* enc_b_ih; /* encoder.rnn.bias_ih_l0              [3H]      */
  const float* enc_b_hh; /* encoder.rnn.bias_hh_l0              [3H]      */
  const float* emb;      /* decoder.embedding.weight            [Vc, E]   */
  const float* dec_w_ih; /* decoder.rnn.weight_ih_l0            [3H, H+E] */
  const float* dec_w_hh; /* decoder.rnn.weight_hh_l0            [3H, H]   */
  const float* dec_b_ih; /* decoder.rnn.bias_ih_l0              [3H]      */
  const float* dec_b_hh; /* decoder.rnn.bias_hh_l0              [3H]      */
  const float* att_wk;   /* decoder.attention.key_layer.weight   [H, H]   */
  const float* att_wq;   /* decoder.attention.query_layer.weight [H, H]   */
  const float* att_v;    /* decoder.attention.energy_layer.weight [1, H]  */
  const float* out_w;    /* decoder.pred_linear.1.weight        [Vc, H]   */
  const float* out_b;    /* decoder.pred_linear.1.bias          [Vc]      */
} PvcrS2vtAttParams;

/* Gradients, same shapes; every non-NULL buffer is overwritten. out_w / out_b are produced by pvcr_vocab_*_bwd. */
typedef struct {
  float *enc_w_ih, *enc_w_hh, *enc_b_ih, *enc_b_hh, *emb, *dec_w_ih, *dec_w_hh, *dec_b_ih, *dec_b_hh, *att_wk,
      *att_wq, *att_v, *out_w, *out_b;
} PvcrS2vtAttGrads;

/* Encoder + attention decoder, teacher-forced (S2VTAttModel.forward in training mode up to, but excluding,
 * the vocabulary projection; model/S2VTAttModel.py:80-96,150-196).
 *   vid_feats [B,N,V]; frame_scale [B,N] or NULL (RationaleNet: sel = vid_feats * p1, model/RationaleNet.py:52);
 *   s_in [B,L] decoder input words (= [<sos>, s[:, :L-1]]);  hs [B,L,H] decoder hidden states (output);
 *   alphas [L,B,N] attention weights (output, may be NULL).
 * Activations needed by the backward pass stay in `workspace`, which must be passed unchanged to _bwd. */
size_t pvcr_s2vtatt_workspace(const PvcrDims* d, int need_frame_grad);
int pvcr_s2vtatt_fwd(const PvcrDims* d, const PvcrS2vtAttParams* p, const float* vid_feats, const float* frame_scale,
                     const int64_t* s_in, float* hs, float* alphas, void* workspace, size_t workspace_bytes,
                     void* stream);
/* Backward of the above given d_hs [B,L,H].  Writes every gradient in g except out_w / out_b; if
 * d_frame_scale != NULL also writes d loss / d frame_scale [B,N] (requires need_frame_grad at sizing). */
int pvcr_s2vtatt_bwd(const PvcrDims* d, const PvcrS2vtAttParams* p, const float* vid_feats, const float* frame_scale,
                     const int64_t* s_in, const float* hs, const float* d_hs, PvcrS2vtAttGrads* g,
                     float* d_frame_scale, void* workspace, size_t workspace_bytes, void* stream);

/* The same backward in two halves for data-parallel callers: part 1 = decoder half (on return att_*, dec_*, emb
 * gradients are final), part 2 = encoder half (enc_* gradients, d_frame_scale), to be called in this order on the same
 * workspace; part 0 = both.  Lets the decoder gradients' all-reduce overlap the encoder sweep. */
int pvcr_s2vtatt_bwd_part(const PvcrDims* d, const PvcrS2vtAttParams* p, const float* vid_feats, const float* frame_scale,
                          const int64_t* s_in, const float* hs, const float* d_hs, PvcrS2vtAttGrads* g,
                          float* d_frame_scale, void* workspace, size_t workspace_bytes, void* stream, int part);

/* Fixed-length beam search over the decoder step.  The reference has no beam search: the search
 * is DEFINED (oracle/captioning_oracle.py: s2vtatt_beam_search) as the plain search over Decoder.forward_step
 * (model/S2VTAttModel.py:125-148) that extends every hypothesis for exactly L steps like the reference's greedy eval
 * branch (:172-191), scores by the sum of log-softmax and keeps the `beam` best of a video's beam x Vc candidates
 * (ties: lower beam * Vc + word first) - so that beam = 1 is the reference's greedy decoding.
 *   ids [B, beam, L] (best hypothesis first), scores [B, beam] (nullable); beam <= 8; nsplit = 3 for fp32-equivalence. */
size_t pvcr_s2vtatt_beam_workspace(const PvcrDims* d, int beam);
int pvcr_s2vtatt_beam(const PvcrDims* d, const PvcrS2vtAttParams* p, const float* vid_feats, const float* frame_scale,
                      int64_t sos_id, int beam, int64_t* ids, float* scores, void* workspace, size_t workspace_bytes,
                      void* stream);

/* S2VTModel parameters (reference state_dict names, model/S2VTModel.py:36-49). */
typedef struct {
  const float* emb;       /* embedding.0.weight   [Vc, E]   */
  const float* rnn1_w_ih; /* rnn1.weight_ih_l0    [3H, V]   */
  const float* rnn1_w_hh; /* rnn1.weight_hh_l0    [3H, H]   */
  const float* rnn1_b_ih; /* rnn1.bias_ih_l0      [3H]      */
  const float* rnn1_b_hh; /* rnn1.bias_hh_l0      [3H]      */
  const float* rnn2_w_ih; /* rnn2.weight_ih_l0    [3H, H+E] */
  const float* rnn2_w_hh; /* rnn2.weight_hh_l0    [3H, H]   */
  const float* rnn2_b_ih; /* rnn2.bias_ih_l0      [3H]      */
  const float* rnn2_b_hh; /* rnn2.bias_hh_l0      [3H]      */
  const float* out_w;     /* linear.1.weight      [Vc, H]   */
  const float* out_b;     /* linear.1.bias        [Vc]      */
} PvcrS2vtParams;
typedef struct {
  float *emb, *rnn1_w_ih, *rnn1_w_hh, *rnn1_b_ih, *rnn1_b_hh, *rnn2_w_ih, *rnn2_w_hh, *rnn2_b_ih, *rnn2_b_hh, *out_w,
      *out_b;
} PvcrS2vtGrads;

/* S2VTModel.forward with the fed-back words given (model/S2VTModel.py:74-145): encode with rnn1/rnn2, then L
 * decoding steps whose input words are s_in [B,L] (teacher forcing: [<sos>, s[:, :L-1]]; with scheduled sampling
 * the caller first obtains the sampled words from pvcr_s2vt_decode_steps).  hs [B,L,H] = rnn2 decoding states
 * (the input of the vocabulary projection).  dims.dropout_p / seed drive the Dropout on the embedded words
 * (embedding.1).  _bwd: given d_hs, writes every gradient in g except out_w / out_b, and d_frame_scale [B,N] if
 * non-NULL.  The workspace carries the activations from _fwd to _bwd. */
size_t pvcr_s2vt_workspace(const PvcrDims* d, int need_frame_grad);
int pvcr_s2vt_fwd(const PvcrDims* d, const PvcrS2vtParams* p, const float* vid_feats, const float* frame_scale,
                  const int64_t* s_in, float* hs, void* workspace, size_t workspace_bytes, void* stream);
int pvcr_s2vt_bwd(const PvcrDims* d, const PvcrS2vtParams* p, const float* vid_feats, const float* frame_scale,
                  const int64_t* s_in, float* hs, const float* d_hs, PvcrS2vtGrads* g, float* d_frame_scale,
                  void* workspace, size_t workspace_bytes, void* stream);

/* RationaleNet generator parameters (reference state_dict names under `gen.`, model/RationaleNet.py:26-30). */
typedef struct {
  const float* w_ih;   /* rnn.weight_ih_l0          [4H, V] */
  const float* w_hh;   /* rnn.weight_hh_l0          [4H, H] */
  const float* b_ih;   /* rnn.bias_ih_l0            [4H]    */
  const float* b_hh;   /* rnn.bias_hh_l0            [4H]    */
  const float* w_ih_r; /* rnn.weight_ih_l0_reverse  [4H, V] */
  const float* w_hh_r; /* rnn.weight_hh_l0_reverse  [4H, H] */
  const float* b_ih_r; /* rnn.bias_ih_l0_reverse    [4H]    */
  const float* b_hh_r; /* rnn.bias_hh_l0_reverse    [4H]    */
  const float* lin_w;  /* linear.weight             [2, 2H] */
  const float* lin_b;  /* linear.bias               [2]     */
} PvcrGenParams;
typedef struct {
  float *w_ih, *w_hh, *b_ih, *b_hh, *w_ih_r, *w_hh_r, *b_ih_r, *b_hh_r, *lin_w, *lin_b;
} PvcrGenGrads;

/* Generator.forward (model/RationaleNet.py:32-54) without materialising sel_vid_feats: writes probs [B,N,2],
 * p1 [B,N] = probs[:,:,1] (pass it as `frame_scale` to pvcr_s2vt(att)_fwd, which computes on vid_feats * p1), and
 * pen [2] = { calc_brevity_loss(probs), calc_cont_loss(probs) } (train_utils.py:73-95).  noise: [B*N,2] Exp(1) draws
 * of F.gumbel_softmax (row b*N+n) or NULL to draw them in-kernel from dims->seed; hard != 0 selects the
 * straight-through one-hot of eval mode.  dims: B, N, V, H, nsplit, dropout_p (Dropout on the LSTM outputs), seed.
 * _bwd: d_p1 [B,N] (the d_frame_scale output of the caption network's _bwd), d_probs [B,N,2] and g_pen (device [2],
 * d loss / d pen) are each optional; writes every gradient in g (w_ih and w_ih_r may be the two halves of one
 * contiguous [8H,V] buffer, which saves a GEMM). */
size_t pvcr_generator_workspace(const PvcrDims* d);
int pvcr_generator_fwd(const PvcrDims* d, const PvcrGenParams* p, const float* vid_feats, const float* noise, float tau,
                       int hard, float* probs, float* p1, float* pen, void* workspace, size_t workspace_bytes,
                       void* stream);
int pvcr_generator_bwd(const PvcrDims* d, const PvcrGenParams* p, const float* vid_feats, float tau, const float* d_p1,
                       const float* d_probs, const float* g_pen, PvcrGenGrads* g, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Step-wise decoding with word feedback.
 * pvcr_s2vtatt_greedy: eval branch of S2VTAttModel (model/S2VTAttModel.py:172-191): fixed L steps, arg-max fed back,
 *   no early stop.  ids [B,L] int64; logits [B,L,Vc] (NULL: not materialised); alphas [L,B,N] (NULL ok).
 *   p->out_w / p->out_b are used.  dims->nsplit = 3 reproduces fp32 arithmetic (token ids match the reference).
 * pvcr_s2vt_decode_steps: eval branch of S2VTModel (model/S2VTModel.py:147-177) when teacher_mask == NULL, and
 *   its scheduled-sampling training branch (:121-141) otherwise: teacher_mask is a HOST array of L ints, entry i
 *   being the outcome of the reference's per-step coin `random.random() < teacher_force_prob`; the word fed to step
 *   i+1 is teacher_words[:, i+1] if teacher_mask[i] else argmax(logits_i).  fed [B,L] receives the fed words
 *   (NULL ok).  dims->dropout_p / seed and out_dropout_p reproduce the masks of pvcr_s2vt_fwd / pvcr_vocab_ce_fwd. */
size_t pvcr_s2vtatt_greedy_workspace(const PvcrDims* d);
int pvcr_s2vtatt_greedy(const PvcrDims* d, const PvcrS2vtAttParams* p, const float* vid_feats, const float* frame_scale,
                        int64_t sos_id, int64_t* ids, float* logits, float* alphas, void* workspace,
                        size_t workspace_bytes, void* stream);
size_t pvcr_s2vt_decode_steps_workspace(const PvcrDims* d);
int pvcr_s2vt_decode_steps(const PvcrDims* d, const PvcrS2vtParams* p, const float* vid_feats, const float* frame_scale,
                           int64_t sos_id, const int64_t* teacher_words, const int32_t* teacher_mask,
                           float out_dropout_p, int64_t* ids, int64_t* fed, float* logits, void* workspace,
                           size_t workspace_bytes, void* stream);

/* Vocabulary projection fused with the loss contract:  logits = Dropout(hs) out_w^T + out_b
 * (model/S2VTAttModel.py:145, model/S2VTModel.py:130), then calc_masked_loss / calc_masked_accuracy /
 * argmax (train_utils.py:37-71, train.py:38).
 *   hs [B*L,H] (row b*L+l); target [B*L] int64; s_len [B] int64 (1 <= s_len <= L).
 *   loss3 [3] = { mean_b( sum_l nll*mask / s_len ), #correct under mask, #mask };  pred [B*L] int64 argmax
 *   (first max index);  lse [B*L] log-sum-exp per token;  token_nll [B*L] (optional) the unmasked per-token loss
 *   lse - logit[target], i.e. criterion(logits, target) of train_utils.py:47-48.  logits_out (optional, ld elements) receives the
 *   fp32 logits for callers that need the reference's logits tensor; pass target = NULL to only project.
 * With nsplit = 1 and logits_out = NULL the logits are never materialised: the GEMM epilogue reduces them per tile
 * (forward) and re-creates them to emit bf16 d logits (backward).
 * _bwd needs the workspace of the matching _fwd untouched; gscale is a device scalar d total / d loss (NULL = 1). */
size_t pvcr_vocab_ce_workspace(int B, int L, int H, int Vc, int nsplit, float dropout_p);
/* Optional: stage out_w for a coming pvcr_vocab_ce_fwd on the same workspace on a side lane (see pvcr_side_mode), so
 * the cast overlaps whatever the caller enqueues in between (the encoder / decoder sweeps).  out_w must not change
 * until that _fwd; a no-op when the side lanes are off or the fused path does not apply. */
int pvcr_vocab_ce_prepare(const float* out_w, int B, int L, int H, int Vc, int nsplit, void* workspace,
                          size_t workspace_bytes, void* stream);
int pvcr_vocab_ce_fwd(const float* hs, const float* out_w, const float* out_b, const int64_t* target,
                      const int64_t* s_len, int B, int L, int H, int Vc, int nsplit, float dropout_p, uint64_t seed,
                      float* loss3, int64_t* pred, float* lse, float* token_nll, float* logits_out,
                      int64_t ld_logits_out, void* workspace, size_t workspace_bytes, void* stream);
int pvcr_vocab_ce_bwd(const float* hs, const float* out_w, const float* out_b, const int64_t* target,
                      const int64_t* s_len, int B, int L,
                      int H, int Vc, int nsplit, float dropout_p, uint64_t seed, const float* gscale, float* d_hs,
                      float* d_out_w, float* d_out_b, float* lse, int64_t* pred, void* workspace,
                      size_t workspace_bytes, void* stream);

/* The decoder halves on caller-given encoder outputs -- `decode(encoder_outs, encoder_final, s)` of both caption nets
 * (model/S2VTAttModel.py:231-243, model/S2VTModel.py:88-145), which is how the reference's SpatialNet drives a caption
 * net after its own per-frame encoder loop (model/SpatialNet.py:140).  enc_outs / out1 are [B,N,H] (row b*N + n; the
 * Python wrapper transposes the reference's [N,B,H]), enc_final / state1 [B,H]; hs [B,L,H] feeds pvcr_vocab_ce_fwd as
 * usual.  Workspaces: pvcr_s2vtatt_workspace(d, 0) / pvcr_s2vt_workspace(d, 0) with any d->V >= 1.  The _bwd calls need
 * the workspace of the matching _fwd untouched; they write every gradient of `grads` except the encoder-GRU entries
 * (enc_* resp. rnn1_w_ih), which may be NULL, plus the gradients on the given encoder outputs and state. */
int pvcr_s2vtatt_decode_fwd(const PvcrDims* d, const PvcrS2vtAttParams* p, const float* enc_outs, const float* enc_final,
                            const int64_t* s_in, float* hs, float* alphas, void* workspace, size_t workspace_bytes,
                            void* stream);
int pvcr_s2vtatt_decode_bwd(const PvcrDims* d, const PvcrS2vtAttParams* p, const int64_t* s_in, const float* hs,
                            const float* d_hs, PvcrS2vtAttGrads* grads, float* d_enc_outs, float* d_enc_final,
                            void* workspace, size_t workspace_bytes, void* stream);
int pvcr_s2vt_decode_fwd(const PvcrDims* d, const PvcrS2vtParams* p, const float* out1, const float* state1,
                         const int64_t* s_in, float* hs, void* workspace, size_t workspace_bytes, void* stream);
int pvcr_s2vt_decode_bwd(const PvcrDims* d, const PvcrS2vtParams* p, const int64_t* s_in, float* hs, const float* d_hs,
                         PvcrS2vtGrads* grads, float* d_out1, float* d_state1, void* workspace, size_t workspace_bytes,
                         void* stream);

/* decode() in eval mode: fixed-length greedy decoding from caller-given encoder outputs (the eval branches of
 * model/S2VTAttModel.py:172-191 and model/S2VTModel.py:147-177 reached through `decode`).  Workspaces:
 * pvcr_s2vtatt_greedy_workspace / pvcr_s2vt_decode_steps_workspace; outputs as pvcr_s2vtatt_greedy / pvcr_s2vt_decode_steps. */
int pvcr_s2vtatt_decode_greedy(const PvcrDims* d, const PvcrS2vtAttParams* p, const float* enc_outs, const float* enc_final,
                               int64_t sos_id, int64_t* ids, float* logits, float* alphas, void* workspace,
                               size_t workspace_bytes, void* stream);
/* The same decoder with the parameter-only work separated from the per-batch work (eval loops decode batch after batch
 * with fixed parameters: eval.py / eval_attention.py call model(vid_feats, None) once per batch).  Exactly one of
 * vid_feats / (enc_outs, enc_final) is given.  Without PVCR_DECODE_REUSE_PREPARED the call first stages the weight planes
 * and the word table  T[w] = W_e Emb[w] + b_ih  (the decoder input projection of every vocabulary word; a decoding step
 * reads row T[argmax]) into the workspace; with it the call trusts that the SAME workspace was prepared by an earlier call
 * with the same dims and parameter values (and not written by anybody else since) and skips that work. */
#define PVCR_DECODE_REUSE_PREPARED 1
int pvcr_s2vtatt_greedy_ex(const PvcrDims* d, const PvcrS2vtAttParams* p, const float* vid_feats, const float* frame_scale,
                           const float* enc_outs, const float* enc_final, int64_t sos_id, int64_t* ids, float* logits,
                           float* alphas, void* workspace, size_t workspace_bytes, int flags, void* stream);
int pvcr_s2vt_decode_greedy(const PvcrDims* d, const PvcrS2vtParams* p, const float* out1, const float* state1,
                            int64_t sos_id, int64_t* ids, float* logits, void* workspace, size_t workspace_bytes,
                            void* stream);

/* One GRU step  h' = GRU(x, h_prev)  -- `encode_step(vid_feat, rnn_state)` of both caption nets
 * (model/S2VTAttModel.py:63-78,219-229, model/S2VTModel.py:57-72: `self.rnn(vid_feat.unsqueeze(0), rnn_state)`), the
 * call SpatialNet makes once per frame (model/SpatialNet.py:127).  x [B,V], h_prev [B,H] or NULL (zeros), torch.nn.GRU
 * parameter layout (gate order r,z,n).  saved [4*B*H] receives r, z, n, W_hn h + b_hn for the backward.
 * _bwd: d_x / d_h_prev overwritten (NULL ok); d_w_ih [3H,V], d_w_hh [3H,H], d_b_ih, d_b_hh [3H] overwritten, or added
 * to when accumulate != 0 (a caller looping over frames accumulates the parameter gradients across its steps). */
size_t pvcr_gru_step_workspace(int B, int V, int H, int nsplit);
int pvcr_gru_step_fwd(const float* x, const float* h_prev, const float* w_ih, const float* w_hh, const float* b_ih,
                      const float* b_hh, int B, int V, int H, int nsplit, float* h_out, float* saved, void* workspace,
                      size_t workspace_bytes, void* stream);
int pvcr_gru_step_bwd(const float* d_h, const float* x, const float* h_prev, const float* w_ih, const float* w_hh,
                      const float* saved, int B, int V, int H, int nsplit, float* d_x, float* d_h_prev, float* d_w_ih,
                      float* d_w_hh, float* d_b_ih, float* d_b_hh, int accumulate, void* workspace,
                      size_t workspace_bytes, void* stream);

/* SpatialNet front (model/SpatialNet.py:76-86,106): conv_feats = ReLU(BN(Conv3x3(ReLU(BN(Conv3x3(x)))))) on the K x K grid
 * features of every frame, x = vid_feats.view(I = B*N, F, K, K) fp32 channels-first as the reference hands it over.
 * Convolution = nine tcgen05 GEMMs on row-shifted views of one flat zero-padded channels-last matrix (csrc/conv.cu); BatchNorm2d
 * with torch's defaults (training: batch statistics, running estimates updated in place with `momentum`; eval: running
 * estimates).  Outputs: conv_feats [I*K*K, H] channels-last (row i*K*K + y*K + x: the keys of the spatial attention) and,
 * if non-NULL, feats_cl [I*K*K, F] = the input features in the same row order (its values).  _bwd: gradients of the eight
 * parameter tensors from d_conv_feats (the input features need none); needs the workspace of the matching _fwd untouched. */
typedef struct {
  const float *conv1_w, *conv1_b;   /* [H, F, 3, 3], [H]   state_dict: conv.0.weight / .bias */
  const float *bn1_w, *bn1_b;       /* [H], [H]            conv.1.weight / .bias */
  const float *conv2_w, *conv2_b;   /* [H, H, 3, 3], [H]   conv.3.weight / .bias */
  const float *bn2_w, *bn2_b;       /* [H], [H]            conv.4.weight / .bias */
} PvcrSpatialFrontParams;
size_t pvcr_spatial_front_workspace(int I, int K, int F, int H, int nsplit);
int pvcr_spatial_front_fwd(int I, int K, int F, int H, int nsplit, const float* vid_feats, const PvcrSpatialFrontParams* p,
                           float* bn1_running_mean, float* bn1_running_var, float* bn2_running_mean, float* bn2_running_var,
                           int training, float eps, float momentum, float* conv_feats, float* feats_cl, void* workspace,
                           size_t workspace_bytes, void* stream);
int pvcr_spatial_front_bwd(int I, int K, int F, int H, int nsplit, const PvcrSpatialFrontParams* p, int training,
                           const float* d_conv_feats, PvcrSpatialFrontParams* grads, void* workspace, size_t workspace_bytes,
                           void* stream);

/* SpatialNet's whole frame loop in one call per direction (model/SpatialNet.py:114-138): for every frame t
 *   q = query_layer(h_{t-1});  alpha_t, ctx_t = attention over the K*K cells (:27-53; keys proj_key = key_layer(conv_feats), values
 *   feats = the frame's input features);  h_t = GRU(ctx_t, h_{t-1}) = caption_net.encode_step (:127), h_{-1} = 0 (:114).
 * proj_key [B,N,Kc,H] and feats [B,N,Kc,F] fp32 as pvcr_spatial_front_fwd / key_layer leave them; w_q [H,H] = query_layer.weight,
 * v [H] = energy_layer.weight, w_ih [3H,F] / w_hh [3H,H] / b_ih / b_hh = the encoder GRU (torch.nn.GRU layout).  Outputs:
 * outs [N,B,H] (torch.cat of the per-frame outputs, :129-132; the final state is its last frame), alphas [N,B,Kc].  The weights are
 * staged once per call, q and W_hh h come from one stacked product, the parameter gradients are products over all frames.
 * _bwd takes the forward call's workspace (unchanged in between), its outs / alphas, and d_outs [N,B,H]; it overwrites d_proj_key
 * [B,N,Kc,H] and the six parameter gradients.  The features need no gradient. */
size_t pvcr_spatial_encode_workspace(int B, int N, int Kc, int H, int F, int nsplit);
int pvcr_spatial_encode_fwd(int B, int N, int Kc, int H, int F, int nsplit, const float* proj_key, const float* feats, const float* w_q,
                            const float* v, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, float* outs,
                            float* alphas, void* workspace, size_t workspace_bytes, void* stream);
int pvcr_spatial_encode_bwd(int B, int N, int Kc, int H, int F, int nsplit, const float* proj_key, const float* feats, const float* w_q,
                            const float* v, const float* w_ih, const float* w_hh, const float* outs, const float* alphas,
                            const float* d_outs, float* d_proj_key, float* d_w_q, float* d_v, float* d_w_ih, float* d_w_hh,
                            float* d_b_ih, float* d_b_hh, void* workspace, size_t workspace_bytes, void* stream);

/* One step of SpatialNet's per-frame attention over the K*K cells (model/SpatialNet.py:27-53, called at :124):
 *   scores[b,c] = v . tanh(q[b] + proj_key[b,c]);  alpha = softmax_c(scores);  ctx[b] = sum_c alpha[b,c] feats[b,c]
 * q [B,H] (row stride q_ld) = query_layer(encoder state); proj_key [B,Kc,H] and feats [B,Kc,Fv] with explicit batch strides
 * (a frame's slice of a [B,N,Kc,.] tensor); alpha [B,Kc], ctx [B,Fv] out.  _bwd: dq [B,H], dproj_key [B,Kc,H] (contiguous) and
 * dv_part [B,H] (sum over b = gradient of energy_layer.weight) from dctx [B,Fv]; the features need no gradient. */
int pvcr_spatial_attn_fwd(int B, int Kc, int H, int Fv, const float* q, int64_t q_ld, const float* proj_key, int64_t pk_batch_stride,
                          const float* feats, int64_t feats_batch_stride, const float* v, float* alpha, float* ctx, void* stream);
int pvcr_spatial_attn_bwd(int B, int Kc, int H, int Fv, const float* dctx, const float* q, int64_t q_ld, const float* proj_key,
                          int64_t pk_batch_stride, const float* feats, int64_t feats_batch_stride, const float* v,
                          const float* alpha, float* dq, float* dproj_key, float* dv_part, void* stream);

/* y[i] = x[i] * mask_i / (1 - p): the mask pvcr_vocab_ce_fwd / _bwd draw for Dropout(hs) under (dropout_p, seed), i the
 * flat index into the [B*L, H] hidden-state matrix (in place allowed).  Replaces nn.Dropout of `pred_linear` / `linear`
 * (model/S2VTAttModel.py:121-122, model/S2VTModel.py:47-49) for callers that take the materialised-logits route, and lets
 * the parity tests hand the very same mask to the reference. */
int pvcr_out_dropout_apply(const float* x, float* y, int64_t n, float dropout_p, uint64_t seed, void* stream);
/* Test hook: minmax[0] / minmax[1] = smallest / largest uniform the in-kernel Philox generator (dropout masks, Gumbel
 * noise of F.gumbel_softmax's exponential_(), model/RationaleNet.py:49) produces over indices [idx0, idx0 + n). */
int pvcr_debug_philox_minmax(uint64_t seed, uint64_t idx0, uint64_t n, float* minmax, void* stream);

/* The loss contract on a materialised logits tensor (callers that use the reference's module API and then its
 * train_utils functions).  pvcr_masked_ce: calc_masked_loss / calc_masked_accuracy / argmax (train_utils.py:37-71,
 * train.py:38) on logits [B*L, Vc] (row stride ld).  Writes loss3 (as pvcr_vocab_ce_fwd), pred, lse, nll [B*L]; if
 * dlogits != NULL also dlogits = d loss / d logits * gscale[0] (gscale NULL = 1; in-place on logits is allowed).
 * pvcr_rationale_penalties(_bwd): calc_brevity_loss / calc_cont_loss (train_utils.py:73-95) on probs [B,N,2] and their
 * gradient for g_pen = device [2] upstream gradients. */
int pvcr_masked_ce(const float* logits, int64_t ld, int B, int L, int Vc, const int64_t* target, const int64_t* s_len,
                   const float* gscale, float* loss3, int64_t* pred, float* lse, float* nll, float* dlogits,
                   int64_t ld_d, void* stream);
int pvcr_rationale_penalties(const float* probs, int B, int N, float* pen, void* stream);
int pvcr_rationale_penalties_bwd(const float* probs, int B, int N, const float* g_pen, float* dprobs, void* stream);

/* Optimizer step of the reference loop (train.py:104-105,157-160): clip_grad_norm_(params, max_norm) followed by
 * torch.optim.Adam(lr, betas, eps, weight_decay).step() (L2-style decay: grad += weight_decay * param), fused into
 * two multi-tensor kernels without host synchronisation.  All tables live in device memory:
 *   tensors[n]        param / grad / exp_avg / exp_avg_sq pointers and element count of every parameter;
 *   chunk_tensor[c], chunk_off[c]   chunk c covers elements [off, off + chunk_elems) of tensor chunk_tensor[c];
 *   partial[n_chunks] scratch; norm_out (nullable) receives the total gradient norm before clipping.
 * max_norm <= 0 disables clipping.  step_dev (nullable): device step counter, incremented by this call and used for
 * the bias corrections (CUDA-graph replays advance it); otherwise step_host (>= 1, already incremented) is used. */
typedef struct {
  float* param;
  float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t numel;
} PvcrAdamTensor;
int pvcr_adam_clip_step(const PvcrAdamTensor* tensors, const int32_t* chunk_tensor, const int64_t* chunk_off,
                        int n_chunks, int chunk_elems, float lr, float beta1, float beta2, float eps,
                        float weight_decay, float max_norm, int64_t* step_dev, int64_t step_host, float* partial,
                        float* norm_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
