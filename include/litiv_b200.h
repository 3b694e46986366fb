/* litiv_b200 — C ABI of the B200-native change-detection hot path.
 *
 * This is the boundary a `lv::CUDA` specialisation of the reference's background subtractors binds to.
 * The reference declares that slot but leaves it empty ("missing impl"):
 *   modules/video/include/litiv/video/BackgroundSubtractionUtils.hpp:136-138
 *   modules/video/include/litiv/video/BackgroundSubtractorLBSP.hpp:77-80
 *   modules/video/src/BackgroundSubtractorLOBSTER.cpp:400-403
 * and its calling code is already written in apps/changedet/src/main.cpp:236-344.
 *
 * Plain pointers and sizes only; all image buffers are caller-owned host memory unless the name says `_device`.
 * Every function returns 0 on success; on failure the message is available from lvb_last_error() (thread-local),
 * mirroring the text of the reference's lvAssert_ exceptions (utils/defines.hpp.in:109-113).
 * There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef LITIV_B200_H
#define LITIV_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct lvb_context* lvb_handle;

enum { LVB_ALGO_LOBSTER = 0, LVB_ALGO_SUBSENSE = 1, LVB_ALGO_PAWCS = 2 };

/* Constructor parameters of the three algorithms (reference defaults in brackets):
 *   SuBSENSE  BackgroundSubtractorSuBSENSE.hpp:52-57  desc offset [3], min colour dist [30], N [50], required [2], avg window [100], rel [0.333]
 *   LOBSTER   BackgroundSubtractorLOBSTER.hpp:50-55   desc thr [4], colour thr [30], N [35], required [2], lbsp offset [0], rel [0.333]
 *   PAWCS     BackgroundSubtractorPAWCS.hpp:51-55     desc offset [2], min colour dist [20], max local words [50], avg window [100], rel [0.333] */
typedef struct lvb_params {
    float rel_lbsp_threshold;
    int32_t lbsp_threshold_offset;
    int32_t desc_dist_threshold;      /* SuBSENSE/PAWCS: nDescDistThresholdOffset ; LOBSTER: nDescDistThreshold */
    int32_t color_dist_threshold;     /* SuBSENSE/PAWCS: nMinColorDistThreshold  ; LOBSTER: nColorDistThreshold */
    int32_t n_samples;                /* nBGSamples / nMaxNbWords */
    int32_t n_required;               /* nRequiredBGSamples */
    int32_t n_samples_for_moving_avgs;
    int32_t n_global_words;           /* PAWCS */
    int32_t median_blur_kernel_size;  /* BGSLBSP_DEFAULT_MEDIAN_BLUR_KERNEL_SIZE [9] */
} lvb_params;

const char* lvb_last_error(void);
/* number of visible CUDA devices (0 when none / driver missing) */
int lvb_device_count(void);
/* fills the reference's default constructor arguments for `algo` */
int lvb_default_params(int algo, lvb_params* out);

/* constructor of BackgroundSubtractor{LOBSTER,SuBSENSE,PAWCS}_<lv::CUDA>; `seed` keys the Philox stream that replaces libc rand() */
int lvb_create(int algo, const lvb_params* params_or_null, int device, uint64_t seed, lvb_handle* out);
int lvb_destroy(lvb_handle h);

/* IIBackgroundSubtractor::initialize(img, ROI)  (BackgroundSubtractionUtils.hpp:28-30; SuBSENSE.cpp:107-186; LOBSTER.cpp:443-457; PAWCS.cpp:431-557)
 * img: 8UC1/8UC3 rows of `step` bytes; roi: null (all pixels) or 8UC1 {0,255} of the same size, continuous */
int lvb_initialize(lvb_handle h, const uint8_t* img, int width, int height, int channels, size_t step, const uint8_t* roi_or_null);
/* IBackgroundSubtractor::apply(img, fgmask, learningRate) (SuBSENSE.cpp:188-612; LOBSTER.cpp:459-581; PAWCS.cpp:559-1523); img continuous, fgmask W*H bytes */
int lvb_apply(lvb_handle h, const uint8_t* img, uint8_t* fgmask, double learning_rate);
/* asynchronous form (the `apply_cuda` async mode of apps/changedet/src/main.cpp:274-282): up to TWO frames may be in flight per instance;
 * the upload of frame k+1 overlaps the kernels of frame k, masks come back on a third stream. lvb_sync_next() waits for the OLDEST
 * frame in flight and delivers its mask (into the buffer passed to the matching lvb_apply_async); lvb_sync() collects everything. */
int lvb_apply_async(lvb_handle h, const uint8_t* img, uint8_t* fgmask, double learning_rate);
int lvb_sync_next(lvb_handle h);
int lvb_sync(lvb_handle h);
/* one stream, n consecutive frames: lvb_apply_async / lvb_sync_next driven from a C loop (two frames in flight, masks delivered in order
 * into fgmasks[i]); returns when the last mask has landed. The sequence loop of samples/changedet/src/main.cpp:33-74 in one call. */
int lvb_apply_stream(lvb_handle h, const uint8_t* const* imgs, uint8_t* const* fgmasks, int n, const double* learning_rates);
/* n independent streams, one frame each (the lv::WorkerPool pattern of apps/changedet/src/main.cpp:148-154) */
int lvb_apply_batch(lvb_handle* hs, const uint8_t* const* imgs, uint8_t* const* fgmasks, int n, double learning_rate);
/* the same for device-resident frames (rows of d_step bytes; d_fgmasks or its entries may be null): enqueues one frame per
 * instance from a small pool of host threads and returns without waiting (order later work with lvb_flush / lvb_sync).
 * Both batch calls may be issued from several host threads at once (e.g. one per GPU): batches are serialised inside. */
int lvb_apply_batch_device(lvb_handle* hs, const uint8_t* const* d_imgs, size_t d_step, uint8_t* const* d_fgmasks, int n, double learning_rate);
/* device-resident variant: d_img rows of d_step bytes, d_fgmask W*H bytes (or null), asynchronous on the instance's stream */
int lvb_apply_device(lvb_handle h, const uint8_t* d_img, size_t d_step, uint8_t* d_fgmask_or_null, double learning_rate);

/* SuBSENSE runs its mask post-processing on a side stream of the instance (it overlaps the next frame's scan): lvb_flush makes
 * the instance's stream (lvb_stream) wait for everything enqueued so far, so that work or events the caller puts on that stream
 * afterwards are ordered behind the masks written by lvb_apply_device. Host-side calls (lvb_apply, lvb_sync*, state access)
 * already wait for all streams. */
int lvb_flush(lvb_handle h);

/* getBackgroundImage / getBackgroundDescriptorsImage (SuBSENSE.cpp:614-649; LOBSTER.cpp:583-620; PAWCS.cpp:1525-1594) */
int lvb_get_background_image(lvb_handle h, uint8_t* out);
int lvb_get_background_descriptors_image(lvb_handle h, uint16_t* out);
/* the same into DEVICE memory (W*H*C bytes, continuous): the display path of apps/changedet/src/main.cpp:304-307 passes a cv::cuda::GpuMat */
int lvb_get_background_image_device(lvb_handle h, uint8_t* d_out);
/* IIBackgroundSubtractor::validateROI (BackgroundSubtractionUtils.hpp:38, .cpp:28-36): clears every ROI pixel closer than `border` to the frame
 * edge, in place (border = 2 = LBSP::PATCH_SIZE/2 for the three LBSP-based subtractors: BackgroundSubtractorLBSP.hpp, m_nROIBorderSize) */
int lvb_validate_roi(uint8_t* roi, int width, int height, int border);
/* refreshModel(fSamplesRefreshFrac, bForceFGUpdate) (SuBSENSE.cpp:80-105; LOBSTER.cpp:410-441) */
int lvb_refresh_model(lvb_handle h, float frac, int force_fg);
/* BackgroundSubtractorPAWCS::refreshModel(nBaseOccCount, fOccDecrFrac, bForceFGUpdate) (PAWCS.cpp:107-429) */
int lvb_pawcs_refresh_model(lvb_handle h, uint32_t base_occ, float decr_frac, int force_fg);
/* setAutomaticModelReset / getROICopy / setROI (BackgroundSubtractionUtils.cpp:24-52) */
int lvb_set_auto_model_reset(lvb_handle h, int enabled);
int lvb_get_roi(lvb_handle h, uint8_t* out);
int lvb_set_roi(lvb_handle h, const uint8_t* roi);
/* getDefaultLearningRate (SuBSENSE/PAWCS 0, LOBSTER 16) */
double lvb_default_learning_rate(int algo);

/* LBSP::compute2 dense (features2d/src/LBSP.cpp:102-152, 231-237): use_rel=0 -> absolute threshold `thr`;
 * use_rel=1 -> saturate(ref*rel + thr). ref_or_null = LBSP::setReference image. out: [H][W][C] u16, 2-px border untouched. */
int lvb_lbsp_compute(const uint8_t* img, const uint8_t* ref_or_null, int width, int height, int channels,
                     int use_rel, float rel, int thr, uint16_t* out, int device);

/* LBSP::computeDescriptor_gradient<C, nAbsOffset = 20, nRelShift = 2> evaluated densely (features2d/include/litiv/features2d/LBSP.hpp:
 * 235-256; the per-pixel primitive of the LBSP edge detector, imgproc/src/EdgeDetectorLBSP.cpp:253). out: [H][W][4] bytes = gradX (int8),
 * gradY (int8), gradient magnitude (0..16), 0 -- the layout of the detector's gradient map (:196); the 2-px border is (0,0,0,0). */
int lvb_lbsp_gradient(const uint8_t* img, int width, int height, int channels, uint8_t* out, int device);

/* the bit-packed mask operators that replace the reference's OpenCV calls (SuBSENSE.cpp:536-554), standalone on byte masks:
 * op 0 cv::dilate / 1 cv::erode with a (2*param+1)^2 rect (param 1 or 3) ; 2 cv::medianBlur(param) on a binary mask ;
 * 3 cv::floodFill((0,0),255)+bitwise_not ("holes": background not 4-connected to the border, needs src(0,0)==0) */
int lvb_mask_op(int op, const uint8_t* src, uint8_t* dst, int width, int height, int param, int device);

/* CDnet-style evaluation, the step right after apply() in the reference's loop: lv::BinClassif::accumulate(oClassif, oGT, oROI)
 * (modules/datasets/src/metrics.cpp:21-61) computed on the device. counters[6] = TP, TN, FP, FN, SE, DC (BinClassif::CountersList,
 * datasets/include/litiv/datasets/metrics.hpp:40-48) are ADDED to. gt: host W*H CDnet labels (0 negative, 50 shadow, 85 out of
 * scope, 170 unknown, 255 positive) or null (every pixel is a don't-care); roi: host W*H or null (pixels equal to 0 are don't-care).
 * lvb_binclassif_accumulate scores the instance's latest foreground mask where it lives (bit-packed in HBM: no mask read-back);
 * lvb_binclassif scores a caller-provided host mask. lvb_binclassif_metrics = BinClassifMetrics (metrics.hpp:213-257):
 * out[8] = recall, specificity, FPR, FNR, PBC, precision, F-measure, MCC. */
int lvb_binclassif_accumulate(lvb_handle h, const uint8_t* gt_or_null, const uint8_t* roi_or_null, uint64_t counters[6]);
int lvb_binclassif(const uint8_t* classif, const uint8_t* gt_or_null, const uint8_t* roi_or_null, int width, int height, uint64_t counters[6], int device);
int lvb_binclassif_metrics(const uint64_t counters[6], double out[8]);

/* parity / debug: named state buffers in the reference's layout (sample-major [N][H][W][C], maps [H][W]); see DESIGN.md */
int lvb_state_size(lvb_handle h, const char* name, size_t* bytes);
int lvb_state_get(lvb_handle h, const char* name, void* out, size_t bytes);
int lvb_state_set(lvb_handle h, const char* name, const void* in, size_t bytes);
/* instrumentation: out[0..4] = roi_px, samples_scanned, sample_writes, fg_px, frames (accumulated while enabled) */
int lvb_set_collect_stats(lvb_handle h, int enabled);
int lvb_get_stats(lvb_handle h, uint64_t out[5]);
/* number of kernels this library launched since the process started (bench.py's gpu_launches) */
uint64_t lvb_kernel_launch_count(void);
/* per-launch timing of the dominant kernel (SuBSENSE: the scan kernel; LOBSTER / PAWCS: phase A) with CUDA events on the instance's stream: enable, run frames, read+reset */
int lvb_set_profile(lvb_handle h, int enabled);
int lvb_get_profile(lvb_handle h, double* ms_total, uint64_t* launches);
/* same for the second-largest kernel of a SuBSENSE frame (the feedback kernel); zero launches for the other algorithms */
int lvb_get_profile_feedback(lvb_handle h, double* ms_total, uint64_t* launches);
/* same for the two tail passes of the SuBSENSE scan (pixels needing more than two samples), timed together */
int lvb_get_profile_tail(lvb_handle h, double* ms_total, uint64_t* launches);
/* page-locked host buffers: frames / masks living in them are copied straight to / from the device (no staging copy) */
int lvb_host_alloc(void** out, size_t bytes);
int lvb_host_free(void* p);
/* CUDA stream of the instance (cudaStream_t as void*) so callers can time on it with CUDA events */
void* lvb_stream(lvb_handle h);

/* ---- ViBe: BackgroundSubtractorViBe_1ch / _3ch, video/include/litiv/video/BackgroundSubtractorViBe.hpp:50-103,
 * video/src/BackgroundSubtractorViBe.cpp. The reference classes derive from cv::BackgroundSubtractor directly (no ROI, no LBSP
 * layer), hence their own handle type. model_channels = 1 (_1ch: 8UC1 frames only) or 3 (_3ch: 8UC3 frames, or 8UC1 frames expanded
 * like cvtColor(GRAY2BGR), ViBe.cpp:121-124). Constructor defaults: threshold 20, N 20, required 2; apply() default learning rate 16
 * and it must be > 0 (ViBe.hpp:41-47, ViBe.cpp:80). `seed` keys the Philox stream that replaces libc rand(). */
typedef struct lvb_vibe_context* lvb_vibe_handle;
int lvb_vibe_create(int model_channels, int color_dist_threshold, int n_samples, int n_required, int device, uint64_t seed, lvb_vibe_handle* out);
int lvb_vibe_destroy(lvb_vibe_handle h);
/* initialize(oInitImg) (ViBe.cpp:58-76, 115-138): rows of `step` bytes */
int lvb_vibe_initialize(lvb_vibe_handle h, const uint8_t* img, int width, int height, int channels, size_t step);
/* apply(image, fgmask, learningRate) (ViBe.cpp:78-110, 140-191): img continuous, fgmask W*H bytes, synchronous */
int lvb_vibe_apply(lvb_vibe_handle h, const uint8_t* img, int channels, uint8_t* fgmask, double learning_rate);
/* device-resident frame (rows of d_step bytes) and mask (W*H bytes, or null), asynchronous on the instance's stream */
int lvb_vibe_apply_device(lvb_vibe_handle h, const uint8_t* d_img, int channels, size_t d_step, uint8_t* d_fgmask_or_null, double learning_rate);
int lvb_vibe_sync(lvb_vibe_handle h);
/* getBackgroundImage (ViBe.cpp:32-49): W*H*model_channels bytes */
int lvb_vibe_get_background_image(lvb_vibe_handle h, uint8_t* out);
/* parity / checkpointing: the sample model in the reference's layout [N][H][W][C] (m_voBGImg); set != 0 imports it and sets the
 * frame counter that indexes the Philox stream */
int lvb_vibe_model(lvb_vibe_handle h, uint8_t* inout, size_t bytes, int set, uint32_t frame);
/* instrumentation, as for the LBSP-based algorithms: out[0..4] = px, samples_scanned, sample_writes, fg_px, frames */
int lvb_vibe_set_collect_stats(lvb_vibe_handle h, int enabled);
int lvb_vibe_get_stats(lvb_vibe_handle h, uint64_t out[5]);
int lvb_vibe_set_profile(lvb_vibe_handle h, int enabled);
int lvb_vibe_get_profile(lvb_vibe_handle h, double* ms_total, uint64_t* launches);
void* lvb_vibe_stream(lvb_vibe_handle h);

/* ---- PBAS: BackgroundSubtractorPBAS_1ch / _3ch, video/include/litiv/video/BackgroundSubtractorPBAS.hpp:88-147,
 * video/src/BackgroundSubtractorPBAS.cpp (as shipped: self-diffusion on, R2 acceleration off, advanced morphology off). Same shape as the
 * ViBe entry points. Constructor defaults: initial colour distance threshold 30, initial update rate 16, N 35, required 2
 * (PBAS.hpp:47-54); apply()'s learning rate overrides T(x) when > 0 (default -1: use T(x)), PBAS.cpp:177, :411. */
typedef struct lvb_pbas_context* lvb_pbas_handle;
int lvb_pbas_create(int model_channels, int init_color_dist_threshold, float init_update_rate, int n_samples, int n_required, int device,
                    uint64_t seed, lvb_pbas_handle* out);
int lvb_pbas_destroy(lvb_pbas_handle h);
/* initialize(oInitImg) (PBAS.cpp:60-110, 284-326) */
int lvb_pbas_initialize(lvb_pbas_handle h, const uint8_t* img, int width, int height, int channels, size_t step);
/* apply(image, fgmask, learningRateOverride) (PBAS.cpp:112-271, 328-496), synchronous */
int lvb_pbas_apply(lvb_pbas_handle h, const uint8_t* img, int channels, uint8_t* fgmask, double learning_rate_override);
int lvb_pbas_apply_device(lvb_pbas_handle h, const uint8_t* d_img, int channels, size_t d_step, uint8_t* d_fgmask_or_null, double learning_rate_override);
int lvb_pbas_sync(lvb_pbas_handle h);
/* getBackgroundImage (PBAS.cpp:37-54) */
int lvb_pbas_get_background_image(lvb_pbas_handle h, uint8_t* out);
/* parity / checkpointing: "bg_color" / "bg_grad" [N][H][W][C] u8, "R" / "T" / "meanmin" [H][W] f32, "scalars" 2 x f64 (frame counter,
 * m_fFormerMeanGradDist); read-only "rawmask" [H][W] u8 and "lastgrad" [H][W][C] u8 */
int lvb_pbas_state(lvb_pbas_handle h, const char* name, void* inout, size_t bytes, int set);
int lvb_pbas_set_collect_stats(lvb_pbas_handle h, int enabled);
int lvb_pbas_get_stats(lvb_pbas_handle h, uint64_t out[5]);
int lvb_pbas_set_profile(lvb_pbas_handle h, int enabled);
int lvb_pbas_get_profile(lvb_pbas_handle h, double* ms_total, uint64_t* launches);
void* lvb_pbas_stream(lvb_pbas_handle h);

/* EdgeDetectorLBSP (imgproc/include/litiv/imgproc/EdgeDetectorLBSP.hpp:33-83; imgproc/src/EdgeDetectorLBSP.cpp:26-433): the LBSP
 * multi-scale edge detector. create = the constructor (nLevels = 3, dHystLowThrshFactor = 0.5; same argument checks, :31-32).
 * apply_threshold (:391-410): edges[H][W] = 255 on edge pixels for one detection threshold in [0,1] (outside: the default 0.5).
 * apply (:412-433): confidence[H][W] = 16 x the number of thresholds 0/16 .. 15/16 that mark the pixel. img: host, continuous,
 * 8UC1 or 8UC3. Like the reference object the detector keeps its maps between calls (the two mask rows its suppression loop never
 * writes are carried from call to call); get_gradient_map: [H][W][4] = gradX, gradY, magnitude (min over the scales), pad of the
 * latest call. flood_sweeps: relaxation sweeps the hysteresis of the latest call needed (diagnostic). */
typedef struct lvb_edge_context* lvb_edge_handle;
int lvb_edge_create(int levels, double hyst_low_factor, int device, lvb_edge_handle* out);
int lvb_edge_destroy(lvb_edge_handle h);
/* EdgeDetectorLBSP's third constructor argument bNormalizeOutput (EdgeDetectorLBSP.cpp:26-34, 431-432): lvb_edge_apply min-max
 * normalises its confidence map to [0,255] like cv::normalize(NORM_MINMAX). Default off, as in the reference. */
int lvb_edge_set_normalize(lvb_edge_handle h, int normalize_output);
double lvb_edge_default_threshold(void);
int lvb_edge_apply_threshold(lvb_edge_handle h, const uint8_t* img, int width, int height, int channels, uint8_t* edges, double threshold);
int lvb_edge_apply(lvb_edge_handle h, const uint8_t* img, int width, int height, int channels, uint8_t* confidence);
int lvb_edge_get_gradient_map(lvb_edge_handle h, uint8_t* out);
uint64_t lvb_edge_flood_sweeps(lvb_edge_handle h);
/* device-resident variant (frame already in HBM, row pitch d_step; d_edges_or_null: W*H device bytes) and the detector's CUDA stream */
int lvb_edge_apply_threshold_device(lvb_edge_handle h, const uint8_t* d_img, int width, int height, int channels, size_t d_step, uint8_t* d_edges_or_null, double threshold);
void* lvb_edge_stream(lvb_edge_handle h);

#ifdef __cplusplus
}
#endif
#endif
