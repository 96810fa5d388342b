/*
 * jpdse_b200 -- C ABI of the B200-native (sm_100a) hot path of SenseBrain/JPD-SE.
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; every entry point below replaces
 * the ATen/cuDNN dispatch behind one reference call site (cited as file:line relative to the
 * reference tree) and is what a ctypes binding inside `ctu` would bind.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller; nothing is allocated or freed here;
 *   - every function enqueues on `stream` (a cudaStream_t passed as void*) and never synchronises;
 *   - return value 0 = success, negative = error; jpdse_last_error() gives the message of the last
 *     failure on the calling thread;
 *   - activations between kernels are NHWC bf16 ("pixel-major"); `stats` buffers are
 *     double[batch][channels][2] = (sum, sum of squares) accumulated by the conv epilogue and must be
 *     zeroed by the caller before the conv that fills them.
 */
#ifndef JPDSE_B200_H_
#define JPDSE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JPDSE_OK 0
#define JPDSE_ERR_INVALID (-1)
#define JPDSE_ERR_CUDA (-2)
#define JPDSE_ERR_UNSUPPORTED (-3)

/* ABI version of this header (4); bumped on any signature, enum, layout-flag or entry-point change. */
int jpdse_abi_version(void);
/* Message of the last error on this thread ("" if none). Never NULL. */
const char* jpdse_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * Input build: one-hot + instance edges + concat (+ reflect pad, NHWC bf16).
 * Replaces Pix2PixHDModel.preprocess one-hot scatter_ (ctu/models/pix2pixHD_model.py:376-382),
 * get_edges (:774-783), cat(label, edge) (:394) and cat(input_label, image) (:595), and the first
 * ReflectionPad2d(3) of GlobalGenerator (ctu/models/pix2pixHD_networks/networks.py:210).
 *
 *   label     : class ids, (B,1,H,W); label_dtype 0 = float32 (truncated like .long()), 1 = uint8,
 *               2 = int64
 *   instance  : instance ids, (B,1,H,W); inst_dtype 0 = int32, 1 = int16, 2 = int64, 3 = float32
 *   image     : float32 (B,3,H,W) NCHW, already normalised
 *   num_labels: number of one-hot channels (35 on Cityscapes); ids outside [0,num_labels) are an
 *               error in the reference (scatter_ raises); here they set no channel and bump
 *               *bad_label_count (int32 device counter, may be NULL)
 *   out_nhwc  : bf16 (B, H+2*pad, W+2*pad, c_pad) reflect-padded, channel order
 *               [num_labels one-hot, 1 edge, 3 image, zeros up to c_pad]; may be NULL
 *   out_nchw  : float32 (B, num_labels+4, H, W), the reference's `input_concat`; may be NULL
 */
int jpdse_build_input(const void* label, int label_dtype, const void* instance, int inst_dtype,
                      const float* image, int batch, int height, int width, int num_labels,
                      void* out_nhwc, int pad, int c_pad, float* out_nchw, int* bad_label_count,
                      void* stream);

/* Same, from the COMPACT loader output: `image_u8` is uint8 (B,3,H,W) straight from the decoder and the
 * loader's normalisation -- ToTensor (x/255) then Normalize ((x - mean)/std), float32, ctu/data/base_dataset.py -- is
 * fused in, bit-exact with torchvision's; mean / std are HOST arrays of 3 floats (opt.normalize_mean / normalize_std).
 * With uint8 labels and int16 instance ids a 1024x512 image costs 3.1 MB of host->device traffic instead of 10.5 MB. */
int jpdse_build_input_u8(const void* label, int label_dtype, const void* instance, int inst_dtype,
                         const uint8_t* image_u8, const float* mean, const float* std, int batch,
                         int height, int width, int num_labels, void* out_nhwc, int pad, int c_pad,
                         float* out_nchw, int* bad_label_count, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Convolutions as tcgen05/TMEM implicit GEMM fed by TMA.
 * Replace nn.Conv2d / nn.ConvTranspose2d inside GlobalGenerator and ResnetBlock
 * (ctu/models/pix2pixHD_networks/networks.py:210,215,244,246,283-299) and the Binarizer's 1x1 conv
 * (ctu/quantizers/binarize.py:47,51).
 */
enum jpdse_conv_kind {
  JPDSE_CONV3X3_PAD1 = 0,   /* 3x3 stride 1 on an input already padded by 1: x is (B,H+2,W+2,Cin)   */
  JPDSE_CONV3X3_S2 = 1,     /* 3x3 stride 2, zero pad 1: x is (B,H,W,Cin) unpadded, out (B,H/2,W/2)  */
  JPDSE_CONVT3X3_S2 = 2,    /* ConvTranspose 3x3 s2 p1 op1: x is (B,H,W,Cin), out (B,2H,2W,Cout)     */
  JPDSE_CONV7X7_PAD3 = 3,   /* 7x7 stride 1 on an input already padded by 3: x is (B,H+6,W+6,Cin);   */
                            /* Cin*2 bytes must be a multiple of 16 (Cin = 40 for the stem, 64 head) */
  JPDSE_CONV1X1 = 4,        /* 1x1: x is (B,H,W,Cin)                                                 */
  /* data-gradient kinds (backward of the two "PAD" kinds): full correlation with the flipped,       */
  /* channel-transposed filter. x is the output gradient stored with a ZERO border of 2 (6):         */
  /* (B,H+4,W+4,Cin) -> y (B,H+2,W+2,Cout) = gradient w.r.t. the PADDED forward input. The weight    */
  /* handed to jpdse_conv_pack_weights is the FORWARD conv's (Cin_this = Cout_fwd, Cout_this = Cin_fwd) */
  JPDSE_CONV3X3_FULL = 5,
  JPDSE_CONV7X7_FULL = 6,   /* x (B,H+12,W+12,Cin) with Cin*2 bytes a multiple of 16 and 7*Cin <= 64 */
  /* PatchGAN discriminator convs (networks.py:430-449: kernel 4, zero padding 2). x is stored with a ZERO border  */
  /* of 2 (in_pad must be 2): (B,H+4,W+4,Cin); any H, W (the discriminator runs on 257x513, 129x257, 65x129 ...).  */
  JPDSE_CONV4X4_S2 = 7,       /* stride 2: out (B, H/2+1, W/2+1, Cout)                                              */
  JPDSE_CONV4X4_S1 = 8,       /* stride 1: out (B, H+1, W+1, Cout)                                                  */
  /* their data gradients. x = gradient w.r.t. the conv OUTPUT, zero border 2: (B,h+4,w+4,Cin = Cout_fwd), in_h/in_w */
  /* = h, w; y = gradient w.r.t. the forward INPUT, dense (B,out_h,out_w,Cout = Cin_fwd). The weight handed to       */
  /* jpdse_conv_pack_weights is the FORWARD conv's (Cout_fwd, Cin_fwd_real, 4, 4); cin_real = Cout_fwd real channels. */
  JPDSE_CONV4X4_S2_DGRAD = 9, /* four output-phase GEMMs; out_h/out_w (forward input size: 2h-2 or 2h-1) are required */
  JPDSE_CONV4X4_S1_FULL = 10, /* out (B, h-1, w-1, Cout)                                                            */
  /* 3x3 stride 1 on an input padded by 1 with FEW channels (the VGG19's RGB input conv, networks.py:479): x is       */
  /* (B,H+2,W+2,Cin) with Cin*2 bytes a multiple of 16 and 3*Cin <= 64; the 3*Cin contiguous elements under a filter  */
  /* row are one K block (K = 3 x 64 instead of 9 x 64). The buffer must extend 128 B past its last pixel.            */
  JPDSE_CONV3X3_PAD1_NARROW = 11,
  /* ---- ABI version 3 */
  /* JPDSE_CONV3X3_FULL on a gradient stored in the SHARED-BORDER layout (see JPDSE_PAD_SHARED; in_pad must be 2):   */
  /* the whole batch is ONE flat run of positions with row pitch W+2 = the output width, so the GEMM's M dimension   */
  /* is exactly the B*(H+2)*(W+2) output pixels (the per-image form computes (H+2)*(W+4) positions per image and      */
  /* rounds every image up to a tile: 152 tiles instead of 144 for the 1024-channel ResnetBlocks at batch 2 -- the    */
  /* difference between one wave and two on 148 SMs). y is the same dense (B,H+2,W+2,Cout) tensor.                    */
  JPDSE_CONV3X3_FULL_SHARED = 12
};
/* SHARED-BORDER layout of a zero-bordered gradient tensor, selected by OR-ing this flag into the pad argument
 * (pad = p | JPDSE_PAD_SHARED) of jpdse_instnorm_backward_apply / jpdse_instnorm_backward_fused (dx_pad) and
 * jpdse_conv_wgrad (dy_pad, JPDSE_CONV3X3_PAD1 only): rows have a pitch of W+p positions -- the p zero positions behind a
 * row are the right border of that row AND the left border of the next -- and images a stride of (H+p) rows -- the p
 * zero rows behind an image are its bottom border and the next image's top border. Pixel (b,i,j) lives at position
 * b*(H+p)*(W+p) + (i+p)*(W+p) + (j+p); the buffer holds B*(H+p)*(W+p) + p*(W+p) + p positions of C channels (the
 * producers write all of them, zeros included) and should extend 128 further positions (tile over-read, never stored). */
#define JPDSE_PAD_SHARED 0x100
enum jpdse_conv_epilogue {
  JPDSE_EPI_RAW_STATS = 0,      /* y = bf16 NHWC raw conv output, stats += (sum, sumsq) per (b,c)    */
  JPDSE_EPI_BIAS_TANH_NCHW = 1, /* y = float32 NCHW tanh(conv + bias)            (networks.py:246)   */
  JPDSE_EPI_SIGN_NCHW = 2,      /* y = float32 NCHW sign(tanh(conv))             (binarize.py:51-54) */
  JPDSE_EPI_RAW = 3,            /* y = bf16 NHWC, no statistics (gradients)                           */
  JPDSE_EPI_BIAS_ACT = 4,       /* y = bf16 NHWC LeakyReLU_slope(conv + bias) (slope 0 = ReLU): PatchGAN layer 0 */
                                /* (networks.py:430) and the VGG19 convs (:474-504)                    */
  JPDSE_EPI_BIAS_NCHW = 5       /* y = float32 NCHW conv + bias: the PatchGAN's 1-channel output conv (:447) */
};
typedef struct jpdse_conv_desc {
  int kind;      /* enum jpdse_conv_kind */
  int epilogue;  /* enum jpdse_conv_epilogue */
  int batch;
  int in_h, in_w; /* logical (unpadded) input height / width */
  int in_pad;     /* border (pixels) physically present around x: must be 1 / 3 for the PAD1 / PAD3  */
                  /* kinds (it is the conv's padding); for the other kinds it is skipped over        */
  int cin;        /* input channels as stored (multiple of 64, or 40 for the stem) */
  int cin_real;   /* channels of the torch weight (<= cin); extra stored channels multiply zero weights */
  int cout;       /* output channels */
  /* ---- ABI version 2 */
  int out_pad;    /* bf16 NHWC epilogues: y is (B, out_h+2*out_pad, out_w+2*out_pad, cout) and only its interior is  */
                  /* written (the border belongs to the caller: zero for the next zero-padded conv)                 */
  int out_h, out_w; /* explicit output size, only read by JPDSE_CONV4X4_S2_DGRAD (0 elsewhere)                      */
  float slope;    /* JPDSE_EPI_BIAS_ACT: negative slope of the LeakyReLU (0.2 PatchGAN, 0 = ReLU)                     */
  int cout_real;  /* data-gradient kinds 9 / 10: input channels of the FORWARD weight (<= cout; 0 = cout); output     */
                  /* channels beyond it are written as zeros                                                       */
} jpdse_conv_desc;

/* Bytes of the packed (bf16, K-major, tap-blocked) weight buffer for this conv. 0 on error. */
size_t jpdse_conv_packed_weight_bytes(const jpdse_conv_desc* d);
/* Pack a float32 torch-layout weight -- Conv2d (Cout,Cin,kh,kw) or ConvTranspose2d (Cin,Cout,kh,kw)
 * -- into the layout jpdse_conv_forward expects. */
int jpdse_conv_pack_weights(const jpdse_conv_desc* d, const float* w, void* w_packed, void* stream);
/* y = conv(x). `bias` is only read by JPDSE_EPI_BIAS_TANH_NCHW; `stats` only by RAW_STATS. */
int jpdse_conv_forward(const jpdse_conv_desc* d, const void* x, const void* w_packed,
                       const float* bias, void* y, double* stats, void* stream);
/* Kernel launches one jpdse_conv_forward enqueues (4 for the per-phase ConvTranspose path, else 1). 0 on error. */
int jpdse_conv_launch_count(const jpdse_conv_desc* d);
/* FLOPs (2*MAC, algorithmic: real taps and real channels only) of one jpdse_conv_forward. */
double jpdse_conv_flops(const jpdse_conv_desc* d);

/* ---------------------------------------------------------------------------------------------
 * Weight gradient of one convolution (the wgrad part of what autograd runs for `loss_G.backward()`,
 * ctu/trainers/pix2pixHD_trainer.py:69, through networks.py:210,215,244,246,283-299).
 *
 *   d        : the FORWARD conv descriptor
 *   x        : the forward input exactly as jpdse_conv_forward saw it (bf16 NHWC incl. its border)
 *   dy       : bf16 NHWC gradient w.r.t. the raw conv output, (B, out_h + 2*dy_pad, out_w + 2*dy_pad, cout)
 *              with a border of dy_pad pixels that is skipped (the data-gradient kinds want it zero).
 *              7x7 head (JPDSE_EPI_BIAS_TANH_NCHW): dy is the 8-channel tensor written by
 *              jpdse_tanh_backward_nchw, dy_pad must be 6.
 *   dw       : float32 gradient in the torch weight layout -- Conv2d (cout, cin_real, k, k),
 *              ConvTranspose2d (cin, cout, k, k); overwritten, or accumulated into if `accumulate`
 *   workspace: jpdse_conv_wgrad_workspace_bytes(d, dy_pad) bytes of scratch (split-K partial sums)
 */
size_t jpdse_conv_wgrad_workspace_bytes(const jpdse_conv_desc* d, int dy_pad);
int jpdse_conv_wgrad(const jpdse_conv_desc* d, const void* x, const void* dy, int dy_pad, float* dw,
                     int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * InstanceNorm2d(affine=False) backward (networks.py:27-36), fused with the backward of what the
 * forward apply kernel fused: ReLU (networks.py:204), the ResnetBlock skip (:303-305) and
 * ReflectionPad2d (:210,246,275-276,291-292).
 *
 * reduce: dy = [xhat > 0] * (fold(g) + skip), xhat = (raw - mean) * rstd
 *   g     : bf16 (B, H+2*g_pad, W+2*g_pad, C) gradient w.r.t. the reflect-padded layer output; the
 *           border is folded back onto the interior pixels it mirrored
 *   skip  : optional bf16 (B,H,W,C) second gradient of the same tensor (the skip connection)
 *   raw, stats : what the forward conv wrote (raw output + (sum, sumsq))
 *   dy    : bf16 (B,H,W,C) out
 *   sums  : double (B,C,2) += (sum dy, sum dy*xhat); zeroed by the caller
 * apply : dx = rstd * (dy - sum1/n - xhat * sum2/n) -> bf16 (B, H+2*dx_pad, W+2*dx_pad, C), border = 0
 */
int jpdse_instnorm_backward_reduce(const void* g, int g_pad, const void* skip, const void* raw,
                                   const double* stats, void* dy, double* sums, int batch, int height,
                                   int width, int channels, int relu, float eps, void* stream);
int jpdse_instnorm_backward_apply(const void* dy, const void* raw, const double* stats,
                                  const double* sums, void* dx, int dx_pad, int batch, int height,
                                  int width, int channels, float eps, void* stream);
/* Same with a LeakyReLU(slope) mask instead of ReLU (PatchGAN, networks.py:436-445): dy = (xhat > 0 ? 1 : slope) * (...). */
int jpdse_instnorm_backward_reduce_act(const void* g, int g_pad, const void* skip, const void* raw,
                                       const double* stats, void* dy, double* sums, int batch, int height,
                                       int width, int channels, int relu, float slope, float eps, void* stream);
/* reduce + apply in ONE launch for small feature maps (height * width <= 2048, e.g. the 1024-channel bottleneck of a
 * 1024x512 image): one CTA owns all pixels of (image, 8 channels), the sums are a block reduction and dy stays in
 * registers. Same arguments as the pair; dy (optional, dense bf16 (B,H,W,C)) receives the masked gradient when the caller
 * needs it (the ResnetBlock skip connection). Returns JPDSE_ERR_UNSUPPORTED for larger maps. */
int jpdse_instnorm_backward_fused(const void* g, int g_pad, const void* skip, const void* raw, const double* stats,
                                  void* dy, void* dx, int dx_pad, int batch, int height, int width, int channels,
                                  int relu, float slope, float eps, void* stream);
/* Head: nn.Tanh backward (networks.py:246). grad_out / out: float32 NCHW (B,channels<=8,H,W);
 * d_pre: bf16 (B,H+12,W+12,8) = grad_out * (1 - out^2), zero border 6, channels >= `channels` zero;
 * dbias: float32 (channels) += sum over (B,H,W) of d_pre (zeroed by the caller). */
int jpdse_tanh_backward_nchw(const float* grad_out, const float* out, void* d_pre, float* dbias,
                             int batch, int channels, int height, int width, void* stream);

/* ---------------------------------------------------------------------------------------------
 * InstanceNorm apply (+ReLU) (+residual) (+reflect pad) -- the second half of
 * nn.InstanceNorm2d(affine=False, eps=1e-5) (networks.py:27-36) whose statistics were reduced by
 * the producing conv's epilogue; ReLU (networks.py:204), ResnetBlock skip add (:303-305) and the
 * ReflectionPad2d in front of the next conv (:210,246,275-276,291-292) are fused in.
 *
 *   raw      : bf16 (B,H,W,C) raw conv output
 *   stats    : double (B,C,2)
 *   residual : bf16 (B,H+2*pad,W+2*pad,C) or NULL; added after normalisation (no ReLU after the add)
 *   out      : bf16 (B,H+2*pad,W+2*pad,C); pad > 0 => reflect padding of the normalised tensor
 */
int jpdse_instnorm_apply(const void* raw, const double* stats, const void* residual, void* out,
                         int batch, int height, int width, int channels, int pad, int relu,
                         float eps, void* stream);

/* ---------------------------------------------------------------------------------------------
 * PatchGAN discriminator + feature losses of the training step:
 * MultiscaleDiscriminator / NLayerDiscriminator (networks.py:371-471), discriminate (pix2pixHD_model.py:451-460),
 * feature matching (:746-753), VGGLoss / Vgg19 (networks.py:124-139, 474-504). The convolutions are jpdse_conv_forward /
 * jpdse_conv_wgrad with the JPDSE_CONV4X4_* kinds (VGG: JPDSE_CONV3X3_PAD1 on zero-bordered tensors, JPDSE_EPI_BIAS_ACT).
 *
 * d_input: out (B, Ho+2*out_pad, Wo+2*out_pad, c_pad) bf16, interior = cat(a, b) over channels (a: float32 (B,ca,H,W),
 *   b: float32 (B,cb,H,W) or NULL) -- the reference's torch.cat((input_label, image), dim=1) -- channels >= ca+cb zero;
 *   pool = 1 first applies nn.AvgPool2d(3, stride=2, padding=1, count_include_pad=False) (networks.py:387): Ho = (H-1)/2+1.
 *   The border is NOT written (the caller keeps it zero).
 * d_input_backward: out float32 (B,c,H,W) = g0[..., c0:c0+c] + AvgPool backward of g1[..., c0:c0+c]; g0 dense bf16
 *   (B,H,W,c_stored) = gradient w.r.t. the full-resolution input, g1 dense (B,(H-1)/2+1,(W-1)/2+1,c_stored) or NULL.
 */
int jpdse_d_input(const float* a, int ca, const float* b, int cb, void* out, int batch, int height, int width,
                  int c_pad, int pool, int out_pad, void* stream);
int jpdse_d_input_backward(const void* g0, const void* g1, float* out, int batch, int height, int width,
                           int c_stored, int c0, int c, void* stream);
/* The same discriminator input(s) straight from the ids: channels [one-hot(label) (num_labels), instance edge, image (3)],
 * i.e. preprocess (pix2pixHD_model.py:376-396) + cat with the image (:451-460) [+ the AvgPool] without the float32
 * (B,39,H,W) tensor. label / instance and their dtype codes as in jpdse_build_input; image_a / image_b: float32 (B,3,H,W);
 * out_a / out_b: bf16 (B,Ho+2*out_pad,Wo+2*out_pad,c_pad). The second image / output pair is optional (both NULL): the fake
 * and the real pass share the label channels, so both operands are written in one pass. Border not written. */
int jpdse_d_input_ids(const void* label, int label_dtype, const void* instance, int inst_dtype, const float* image_a,
                      void* out_a, const float* image_b, void* out_b, int batch, int height, int width,
                      int num_labels, int c_pad, int pool, int out_pad, void* stream);
/* InstanceNorm2d apply + LeakyReLU(slope): out (B,H+2*out_pad,W+2*out_pad,C) bf16 with a ZERO border (written). */
int jpdse_instnorm_apply_act(const void* raw, const double* stats, void* out, int batch, int height, int width,
                             int channels, int out_pad, float slope, float eps, void* stream);
/* Backward through y = LeakyReLU_slope(conv + bias) (no norm): d_pre = (g + skip) * (f > 0 ? 1 : slope).
 *   g: bf16 (B,H+2*g_pad,W+2*g_pad,C), border skipped (the data gradient of a zero-padded conv also holds the gradient
 *   w.r.t. the padding); skip (optional): dense bf16 (B,H,W,C); f: the stored activation (B,H+2*f_pad,W+2*f_pad,C);
 *   d_pre: bf16 (B,H+2*out_pad,W+2*out_pad,C), zero border written; dbias (optional): float32 (C) += sum of d_pre. */
int jpdse_act_backward(const void* g, int g_pad, const void* skip, const void* f, void* d_pre, float* dbias,
                       int batch, int height, int width, int channels, int f_pad, int out_pad, float slope,
                       void* stream);
/* nn.L1Loss numerator between two bf16 tensors of one stored shape: *sum += sum |a - b| over n_elements (borders and pad
 * channels are zero in both). Backward: out dense bf16 (B,H,W,C) = sign(a - b) * (*scale_dev) * scale_host, a / b stored
 * with a border of `pad` (scale_dev: optional device scalar, e.g. the upstream gradient). */
int jpdse_l1_pair(const void* a, const void* b, size_t n_elements, double* sum, void* stream);
int jpdse_l1_pair_backward(const void* a, const void* b, void* out, const float* scale_dev, float scale_host,
                           int batch, int height, int width, int channels, int pad, void* stream);
/* nn.MaxPool2d(2, 2) (torchvision VGG19): x (B,H+2*in_pad,W+2*in_pad,C) -> y (B,H/2+2*out_pad,W/2+2*out_pad,C), zero border
 * written. Backward: dx dense (B,H,W,C) = g ((B,H/2+2*g_pad,W/2+2*g_pad,C), border skipped) at the first maximum of
 * each window, else 0. */
int jpdse_maxpool2x2(const void* x, void* y, int batch, int height, int width, int channels, int in_pad,
                     int out_pad, void* stream);
int jpdse_maxpool2x2_backward(const void* x, const void* g, int g_pad, void* dx, int batch, int height,
                              int width, int channels, int in_pad, void* stream);
/* ---- ABI version 4
 * The PatchGAN's 1-channel output conv (nn.Conv2d(512, 1, kernel 4, stride 1, padding 2), networks.py:447) as a 1x1 GEMM:
 * z[b][t][p] = sum_c x[b][p][c] * w[t][c] for the 16 taps t over EVERY stored pixel p of the zero-bordered input
 * (JPDSE_CONV1X1 with 16 outputs, JPDSE_EPI_BIAS_NCHW, zero bias: the activation is read once instead of once per tap),
 * then  out[b,0,y,x] = bias + sum_{kh,kw} z[b][kh*4+kw][y+kh][x+kw]  (jpdse_patch_out_gather; z is float32
 * (B,16,H+4,W+4), out float32 (B,1,H+1,W+1)). Backward: dz[b][p][t] = dout[b, py-kh, px-kw] (jpdse_patch_out_scatter:
 * bf16 (B,H+4,W+4,c_pad), channels >= 16 zero) is at once the P operand of the weight gradient and the input of the data
 * gradient, both 1x1 GEMMs. height / width are the conv's INPUT size H, W. */
int jpdse_patch_out_gather(const float* z, const float* bias, float* out, int batch, int height, int width, void* stream);
int jpdse_patch_out_scatter(const float* dout, void* dz, int batch, int height, int width, int c_pad, void* stream);
/* Stored bf16 (B,H+2*pad,W+2*pad,c_stored) -> float32 NCHW (B,channels,H,W) (feature maps handed back to PyTorch). */
int jpdse_nhwc_pad_to_nchw_f32(const void* x, float* y, int batch, int channels, int height, int width, int pad,
                               int c_stored, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Layout helpers used at the boundary (reference tensors are NCHW float32).
 * nchw->nhwc: y is bf16 (B, H+2*pad_reflect, W+2*pad_reflect, c_pad), channels >= `channels` are 0.
 */
int jpdse_nchw_f32_to_nhwc_bf16(const float* x, void* y, int batch, int channels, int height,
                                int width, int pad_reflect, int c_pad, void* stream);
int jpdse_nhwc_bf16_to_nchw_f32(const void* x, float* y, int batch, int channels, int height,
                                int width, void* stream);

/* ---------------------------------------------------------------------------------------------
 * ctu/quantizers forward passes.
 */
/* RoundedIdentity.forward = torch.round, ties to even (ctu/quantizers/round.py:10-11). */
int jpdse_round_f32(const float* x, float* y, size_t n, void* stream);
/* DifferentiableSign eval = x.sign() (ctu/quantizers/binarize.py:41); sign(0)=0, NaN stays NaN. */
int jpdse_sign_f32(const float* x, float* y, size_t n, void* stream);
/* SoftSignFunction.forward with the uniform noise given: y = +1 if (1-x)/2 <= u else -1
 * (ctu/quantizers/binarize.py:20-24). */
int jpdse_softsign_f32(const float* x, const float* u, float* y, size_t n, void* stream);
/* Binarizer in train() mode (ctu/quantizers/binarize.py:44-65): the 1x1 conv is jpdse_conv_forward (JPDSE_CONV1X1,
 * JPDSE_EPI_RAW -> pre_nhwc, bf16 (B,H,W,C)); forward: tanh_out = tanh(pre), y = +1 if (1 - tanh_out)/2 <= noise else -1,
 * both float32 NCHW (B,C,H,W), noise ~ U[0,1) float32 NCHW supplied by the caller; backward (straight-through sign,
 * :26-28): d_pre_nhwc (bf16 NHWC) = grad_y * (1 - tanh_out^2); the conv's gradients are jpdse_conv_wgrad and a
 * jpdse_conv_forward with the transposed 1x1 weight. */
int jpdse_binarizer_train_forward(const void* pre_nhwc, const float* noise, float* y, float* tanh_out,
                                  int batch, int channels, int height, int width, void* stream);
int jpdse_binarizer_train_backward(const float* grad_y, const float* tanh_out, void* d_pre_nhwc, int batch,
                                   int channels, int height, int width, void* stream);
/* Binary codes exported as bytes: (x+1)/2 for x in {-1,+1} (pix2pixHD_model.py:614, test.py:103-110). */
int jpdse_sign_to_bits_u8(const float* x, uint8_t* y, size_t n, void* stream);
/* S2HVQ (ctu/quantizers/s2h_vq.py): x is (rows, center_size) float32 = x_mtrx flattened over
 * (n, code_len); code_book is (n_center, center_size) float32.
 *   scores  : optional (rows, n_center) squared-L2 scores            (_get_score_mtrx, :72-89)
 *   index   : optional int64 (rows): argmin of scores, first index on ties (_hard_quantize :124)
 *   one_hot : optional (rows, n_center) float32 one-hot of index     (:125-129)
 *   soft    : optional (rows, n_center) softmax(-sigma*scores)       (_soft_quantize :107-108)
 */
int jpdse_s2hvq_encode(const float* x, const float* code_book, size_t rows, int center_size,
                       int n_center, float sigma, float* scores, int64_t* index, float* one_hot,
                       float* soft, void* stream);
/* S2HVQ.decode: argmax over the last dim of code_raw (rows, n_center) -> code_book gather
 * (s2h_vq.py:182-183); out is (rows, center_size); index optional int64 (rows). */
int jpdse_s2hvq_decode(const float* code_raw, const float* code_book, size_t rows, int center_size,
                       int n_center, float* out, int64_t* index, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Eval-metric path (validation / test loops: train.py:63-76, test.py:115-125).
 * tensor2im (ctu/utils/misc.py:64-95): out = uint8(clip((x*std + mean)*255, 0, 255)) in float64, C truncation;
 *   x float32 NCHW (B, channels<=3, H, W); out uint8 (B, H, W, channels) (tensor2im returns HWC images);
 *   mean / std: HOST arrays of `channels` doubles (opt.normalize_mean / opt.normalize_std).
 * distortion: *sum += sum over all elements of |ua - ub| (mode 0, L1Loss) or (ua - ub)^2 (mode 1, MSELoss) where
 *   ua / ub are the tensor2im bytes of a / b (pix2pixHD_model.py:636-641); exact integer; the loss is
 *   sum / (B*channels*H*W). `sum` is a device counter zeroed by the caller. */
int jpdse_tensor2im_u8(const float* x, uint8_t* out, int batch, int channels, int height, int width,
                       const double* mean, const double* std, void* stream);
int jpdse_distortion_u8(const float* a, const float* b, unsigned long long* sum, int batch, int channels,
                        int height, int width, int mode, const double* mean, const double* std,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JPDSE_B200_H_ */
