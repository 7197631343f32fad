/*
 * f5_b200.h — C ABI of libf5b200.so, the B200 (sm_100a) implementation of the f5-tts-mlx hot path.
 *
 * The reference (lucasnewman/f5-tts-mlx) has no FFI/plugin boundary: it is pure Python on MLX, and
 * its "operators" are the mlx.nn / mx.fast calls inside f5_tts_mlx/{cfm,dit,convnext_v2,rope,audio}.py.
 * Each entry point below replaces one of those call sites (cited as file:line of the reference) and
 * is what a binding for that call site would bind.  Conventions:
 *
 *   - every function returns 0 on success or a negative F5_ERR_* code; f5_last_error() returns a
 *     thread-local message for the last failure on the calling thread;
 *   - the caller owns every buffer (weights, activations, workspace); nothing is allocated, no
 *     host synchronisation happens, every launch is ordered on `stream` (a cudaStream_t passed as
 *     void*), so a sequence of calls can be captured into a CUDA graph;
 *   - all pointers are DEVICE pointers unless the parameter name starts with `h_`;
 *   - activations are channels-last (rows = batch*frames, row-major), bf16 operands for the tensor
 *     cores ("bf16" below = __nv_bfloat16 bits), fp32 for the residual stream, statistics, softmax
 *     and ODE state.  Linear weights are (out_features, in_features) row-major bf16, exactly the
 *     reference's nn.Linear.weight layout.
 *   - there is NO CPU fallback: if no sm_100 device is present every compute entry point fails with
 *     F5_ERR_NO_DEVICE.
 */
#ifndef F5_B200_H_
#define F5_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define F5_OK 0
#define F5_ERR_INVALID (-1)   /* bad argument (shape/alignment/null)            */
#define F5_ERR_CUDA (-2)      /* a CUDA runtime / driver call failed             */
#define F5_ERR_NO_DEVICE (-3) /* no sm_100 GPU: this library has no CPU path     */

const char* f5_last_error(void);
/* library/ABI version (major*1000+minor) and the device check used by every entry point */
int f5_abi_version(void);
int f5_device_check(void);

/* ------------------------------------------------------------------------------------------ *
 * Dense / implicit-conv GEMM on tcgen05 tensor cores:  out = epilogue(A · W^T)
 * Replaces every nn.Linear / nn.Conv1d on the path (dit.py:33-38,77,94-99,136-143,170,249,267,
 * 286,399; convnext_v2.py:35-44) together with the elementwise ops the reference applies to its
 * result (bias, GELU/Mish, RoPE rope.py:94-107, "* mask" dit.py:172-173, AdaLN gate + residual
 * dit.py:319,323).
 * ------------------------------------------------------------------------------------------ */
enum { F5_ACT_NONE = 0, F5_ACT_GELU_TANH = 1, F5_ACT_GELU_ERF = 2, F5_ACT_MISH = 3 };

typedef struct f5_gemm_args {
  /* operands */
  const void* a;      /* bf16 [rows, lda]; rows = M (flat) or num_batches*rows_per_batch        */
  int64_t lda;        /* elements, multiple of 8                                                */
  const void* w;      /* bf16 [N, ldw] (out_features major)                                     */
  int64_t ldw;        /* elements, multiple of 8                                                */
  int32_t m, n, k;    /* k: reduction length per tap                                            */
  /* row -> (utterance, frame) mapping */
  int32_t rows_per_batch; /* frames per utterance; 0: single utterance of m rows               */
  int32_t num_batches;    /* >=1                                                                */
  int32_t batched_tiles;  /* 1: tiles never straddle utterances (required for conv_taps > 1)    */
  /* implicit 1-D convolution over frames (conv_taps = 1: plain GEMM).  W is then
   * [N, conv_taps * k_pad] with k_pad = round_up(k, 64), tap-major.                            */
  int32_t conv_taps;
  int32_t conv_pad;
  int32_t conv_grouped;   /* 1: block-diagonal groups of 64 channels (k must be 64)             */
  /* epilogue */
  int32_t act;            /* F5_ACT_*                                                            */
  int32_t out_bf16;       /* 1: bf16 output, 0: fp32                                            */
  const float* bias;      /* [n] or NULL                                                        */
  void* out;              /* [rows, ldo]                                                        */
  int64_t ldo;
  const float* resid;     /* fp32 [rows, ldr] or NULL; may alias out                            */
  int64_t ldr;
  const float* gate;      /* fp32 [num_batches, gate_ld] or NULL                                */
  int64_t gate_ld;
  const int32_t* row_len; /* [num_batches] valid frames (rows beyond are written as 0) or NULL  */
  const float* rope;      /* fp32 [rows_per_batch, 32, 2] (cos,sin) or NULL                     */
  int32_t rope_cols;      /* columns [0, rope_cols) are rotated in adjacent pairs               */
  float q_scale;          /* columns [0, q_cols) are multiplied by q_scale after the rotation   */
  int32_t q_cols;
  int32_t tile_n;         /* 0 = auto, else 64 | 128                                            */
} f5_gemm_args;

int f5_gemm_bf16(const f5_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* F5_B200_H_ */
