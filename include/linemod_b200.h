/*
 * linemod_b200.h -- C ABI of the B200-native LINEMOD matcher (liblinemod_b200.so).
 *
 * Drop-in boundary for the one call the reference ROS package makes into its matcher,
 *     rgbdDetector::linemod_detection -> cv::linemod::Detector::match      /root/reference/src/rgbdDetector.cpp:31-34
 * plus the Detector calls around it (construction, addTemplate, read/readClass, write/writeClass, getTemplates,
 * classIds).  Every entry point below names the reference call site it replaces.  Plain pointers and sizes only;
 * images are borrowed host memory described by lm_image (rows may be strided, exactly like the cropped ROI
 * `mat_rgb(crop)` the service passes, src/linemod_ensenso_detect_3_mult_detect_service.cpp:324-326).
 *
 * All compute runs in hand-written sm_100a CUDA kernels; there is no CPU fallback: lm_create fails with
 * LM_E_CUDA when no CUDA device is usable.
 *
 * Error convention: functions return LM_OK (0) or a negative LM_E_* code; the message of the last failure on the
 * calling thread is lm_last_error().  (The reference's OpenCV raises cv::Exception from CV_Assert at the same
 * conditions; the C++ facade include/linemod_b200.hpp rethrows.)
 *
 * Threading: CUDA streams + workspaces belong to the lm_detector.  Calls on one handle must be serialised by the caller;
 * distinct handles are independent (lm_group below drives one handle per GPU from its own threads).
 */
#ifndef LINEMOD_B200_H_
#define LINEMOD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LM_OK 0
#define LM_E_INVALID (-1)   /* bad argument / violated CV_Assert-style precondition */
#define LM_E_CUDA (-2)      /* CUDA runtime failure or no device */
#define LM_E_IO (-3)        /* file could not be read / written / parsed */
#define LM_E_NOTFOUND (-4)  /* unknown class id / template id */
#define LM_E_STATE (-5)     /* call sequence error (e.g. debug fetch before any match) */

#define LM_MAX_MODALITIES 4
#define LM_MAX_LEVELS 4
#define LM_MAX_FEATURES 63 /* per template, modality and level: u8 accumulation of responses <= 4 stays exact */

/* lm_image.type */
#define LM_8UC3 0  /* BGR, ColorGradient source */
#define LM_16UC1 1 /* depth in mm, DepthNormal source */
#define LM_8UC1 2  /* mask (non-zero = keep) / quantised image */

/* lm_modality_desc.type */
#define LM_COLOR_GRADIENT 0
#define LM_DEPTH_NORMAL 1

typedef struct lm_detector lm_detector;

typedef struct {
  const void* data; /* host memory, caller-owned, borrowed for the duration of the call */
  int32_t rows, cols, type;
  size_t step; /* bytes between rows */
} lm_image;

typedef struct {
  void* data; /* host memory, caller-owned, written by the library */
  int32_t rows, cols, type;
  size_t step;
} lm_image_out;

/* cv::linemod::ColorGradient(weak_threshold, num_features, strong_threshold) /
 * cv::linemod::DepthNormal(distance_threshold, difference_threshold, num_features, extract_threshold);
 * defaults as default-constructed at /root/reference/src/renderer.cpp:180-181. */
typedef struct {
  int32_t type;
  float weak_threshold;   /* CG, default 10 */
  float strong_threshold; /* CG, default 55 */
  int32_t distance_threshold;   /* DN, default 2000 */
  int32_t difference_threshold; /* DN, default 50 */
  int32_t extract_threshold;    /* DN, default 2 */
  int32_t num_features;         /* both, default 63 */
} lm_modality_desc;

/* cv::linemod::Match {x, y, similarity, class_id, template_id}; class_index indexes lm_class_id() (std::map order) */
typedef struct {
  int32_t x, y, template_id, class_index;
  float similarity;
} lm_match_rec;

typedef struct {
  int32_t x, y, width, height;
} lm_rect;

/* cv::linemod::Template header {width, height, pyramid_level} + feature count; features are (x, y, label) int triples */
typedef struct {
  int32_t width, height, pyramid_level, num_features;
} lm_template_hdr;

/* ------------------------------------------------------------------------------------------------ lifecycle */
/* cv::linemod::Detector(modalities, T_pyramid)                       /root/reference/src/renderer.cpp:179-185 */
int lm_create(const int32_t* T, int levels, const lm_modality_desc* mods, int M, lm_detector** out);
/* readLinemod(): Detector::read(fs.root()) + readClass per "classes" entry   src/rgbdDetector.cpp:1668-1680 */
int lm_create_from_yaml(const char* path, lm_detector** out);
/* writeLinemod(): Detector::write + writeClass per class into one file       src/renderer.cpp:56-70 */
int lm_write_yaml(const lm_detector* det, const char* path);
/* Binary template cache: the detector of a templates.yml as flat, checksummed arrays.  The reference's service calls
 * readLinemod() -- a YAML parse of thousands of templates -- on every request
 * (src/linemod_ensenso_detect_3_mult_detect_service.cpp:1784,1851); a cache written once with lm_write_cache loads
 * without parsing.  Same model either way (templates, classes, modalities, T): matching results are identical. */
int lm_create_from_cache(const char* path, lm_detector** out);
int lm_write_cache(const lm_detector* det, const char* path);
/* Detector::readClasses(class_ids, format) / writeClasses(format), format default "templates_%s.yml.gz" */
int lm_read_classes(lm_detector* det, const char* const* class_ids, int n_ids, const char* format);
int lm_write_classes(const lm_detector* det, const char* format);
void lm_destroy(lm_detector* det);
const char* lm_last_error(void);
/* Page-locked host memory for frames: lm_match copies such images to the device asynchronously without staging. */
void* lm_alloc_pinned(size_t bytes);
void lm_free_pinned(void* p);
/* CUDA device the handle computes on (default: current device at lm_create) */
int lm_device(const lm_detector* det);

/* ------------------------------------------------------------------------------------------------ introspection */
int lm_pyramid_levels(const lm_detector* det);                       /* Detector::pyramidLevels */
int lm_get_T(const lm_detector* det, int level);                     /* Detector::getT */
int lm_num_modalities(const lm_detector* det);                       /* Detector::getModalities().size() */
int lm_get_modality(const lm_detector* det, int m, lm_modality_desc* out);
int lm_num_classes(const lm_detector* det);                          /* Detector::numClasses */
int lm_num_templates(const lm_detector* det, const char* class_id);  /* Detector::numTemplates([class_id]); NULL = all */
const char* lm_class_id(const lm_detector* det, int class_index);    /* Detector::classIds()[i]   carmine_detect.cpp:319 */

/* Detector::getTemplates(class_id, template_id)   ..._service.cpp:351,741-744.  hdr receives L*M headers (index
 * l*M+m); feats (nullable) receives all (x,y,label) triples in the same order.  Returns the total feature count. */
int lm_get_templates(const lm_detector* det, const char* class_id, int template_id, lm_template_hdr* hdr, int32_t* feats);

/* ------------------------------------------------------------------------------------------------ training */
/* Detector::addTemplate(sources, class_id, object_mask, &bounding_box)       src/renderer.cpp:308
 * Quantisation runs on the GPU; feature selection (stable sort + scattered selection) on the host.
 * Returns the new template_id (>= 0), -1 when some level lacks candidate features (not an error), < -1 = LM_E_* - 100. */
int lm_add_template(lm_detector* det, const lm_image* sources, int n_sources, const char* class_id,
                    const lm_image* object_mask /*nullable*/, lm_rect* bounding_box /*nullable*/);
/* Host half of addTemplate on quantised maps the caller already holds (what the CUDA front end produces and
 * lm_add_template downloads): quantized[l*M+m] is the level's LM_8UC1 quantisation, magnitudes[l*M+m] the f32
 * gradient magnitude (ColorGradient entries only, tightly packed).  Runs extractTemplate / cropTemplates
 * ([OCV] ColorGradientPyramid::extractTemplate, DepthNormalPyramid::extractTemplate).  No CUDA needed. */
int lm_add_template_from_quantized(lm_detector* det, const lm_image* quantized, const float* const* magnitudes,
                                   const char* class_id, const lm_image* object_mask /*nullable*/,
                                   lm_rect* bounding_box /*nullable*/);
/* Detector::addSyntheticTemplate(templates, class_id).  n_templates must be levels*M.  Returns template_id or LM_E_*. */
int lm_add_synthetic_template(lm_detector* det, const char* class_id, int n_templates, const lm_template_hdr* hdr,
                              const int32_t* feats);

/* cv::linemod::Modality::process(src, mask) -> Ptr<QuantizedPyramid>, and QuantizedPyramid::{quantize, extractTemplate,
 * pyrDown}  ([OCV] linemod.cpp: ColorGradientPyramid / DepthNormalPyramid; the surface behind Detector::addTemplate and
 * Detector::match).  lm_modality_process quantises `src` (CV_8UC3 for ColorGradient, CV_16UC1 for DepthNormal) on
 * the current CUDA device for `levels` pyramid levels at once -- level l is the reference's object after l pyrDown() calls --
 * and keeps the result on the host.  normal_lut: nullable 8000-byte NORMAL_LUT (see lm_set_normal_lut).
 *   lm_qpyramid_quantize(q, l, dst)     quantize(dst): dst caller-allocated CV_8UC1 of lm_qpyramid_size(q, l), masked
 *   lm_qpyramid_extract(q, l, hdr, f)   extractTemplate(templ): 1 = ok, hdr = {-1, -1, l, n} and f (nullable, room for
 *                                       3 * LM_MAX_FEATURES ints) = n (x, y, label) triples in level-l coordinates; 0 = the
 *                                       level lacks candidates (the reference returns false); < 0 = LM_E_*          */
typedef struct lm_qpyramid lm_qpyramid;
int lm_modality_process(const lm_modality_desc* modality, const lm_image* src, const lm_image* mask /*nullable*/, int levels,
                        const uint8_t* normal_lut /*nullable*/, lm_qpyramid** out);
void lm_qpyramid_destroy(lm_qpyramid* q);
int lm_qpyramid_levels(const lm_qpyramid* q);
int lm_qpyramid_size(const lm_qpyramid* q, int level, int* rows, int* cols);
int lm_qpyramid_quantize(const lm_qpyramid* q, int level, lm_image* dst);
int lm_qpyramid_extract(const lm_qpyramid* q, int level, lm_template_hdr* hdr, int32_t* features /*nullable*/);

/* ------------------------------------------------------------------------------------------------ template generation */
/* Training at scale: the loop of /root/reference/src/renderer.cpp:239-329 --
 *     for every view of RendererIterator: render(image, depth, mask, rect); detector->addTemplate(sources, "obj", mask)
 * -- with the renders produced by a CUDA z-buffer rasteriser and addTemplate's feature extraction done on the GPU for a
 * batch of views at a time.  The reference renders with `object_recognition_renderer` (OpenGL + assimp), which is not part
 * of its tree; this library's rasteriser is specified in oracle/render_oracle.cpp (pinhole camera looking at the object
 * origin, principal point at the image centre as K at renderer.cpp:273, depth in u16 millimetres, mask 255). */
typedef struct lm_mesh lm_mesh;
/* A mesh keeps a device copy of its triangles on the GPU of the detector that used it last: share one lm_mesh between
 * detectors on the same device freely, but give concurrent callers on different devices their own.
 * triangles: n_triangles x 3 vertices x (x, y, z) f32, object frame, metres (Renderer3d(stl_file), renderer.cpp:239) */
int lm_mesh_create(const float* triangles, int n_triangles, lm_mesh** out);
int lm_mesh_load_stl(const char* path, lm_mesh** out); /* ASCII or binary STL */
int lm_mesh_num_triangles(const lm_mesh* mesh);
int lm_mesh_get_triangles(const lm_mesh* mesh, float* dst /* n*9 */);
void lm_mesh_destroy(lm_mesh* mesh);

/* Renderer3d::set_parameters(width, height, focal_length_x, focal_length_y, near, far)        renderer.cpp:240-241 */
typedef struct {
  int32_t width, height;
  double fx, fy, near_, far_;
} lm_camera;
/* RendererIterator(renderer, n_points) + angle_step_ / radius_{min,max,step}_               renderer.cpp:242-246.
 * ORK defaults: angle_min -80, angle_max 80. */
typedef struct {
  int32_t n_points, angle_min, angle_max, angle_step;
  float radius_min, radius_max, radius_step;
} lm_view_sphere;
/* RendererIterator::n_templates(), by enumeration (angle innermost, then radius, then sphere point) */
int lm_view_count(const lm_view_sphere* sphere);
/* RendererIterator::view_params for the index-th view: camera position T (object frame) and up vector; optionally the
 * view's radius (D_obj), sphere point and in-plane angle. */
int lm_view_params(const lm_view_sphere* sphere, int index, double T[3], double up[3], float* radius /*nullable*/,
                   int32_t* point_index /*nullable*/, int32_t* angle_deg /*nullable*/);
/* Pose of the object in the camera frame for a view: Pc = R * Po + t (row-major R, OpenCV camera convention) -- what
 * the trainer stores per template (Rs_, Ts_ at renderer.cpp:313-318) in this library's convention. */
int lm_view_pose(const double T[3], const double up[3], double R[9], double t[3]);
/* Renders n_views views (T, up: n_views x 3 each) on the detector's GPU.  Outputs are host buffers, one tightly packed
 * image per view (bgr: rows*cols*3, depth: rows*cols u16, mask: rows*cols u8), each nullable; rects = bounding box of
 * each mask (all zero when the object is not visible). */
int lm_render_views(lm_detector* det, const lm_mesh* mesh, const lm_camera* cam, const double* T, const double* up,
                    int n_views, uint8_t* bgr, uint16_t* depth, uint8_t* mask, lm_rect* rects);
/* Detector::addTemplate for a batch of views (sources[v * n_sources + m], masks[v]; a mask is required): quantisation,
 * candidate extraction, the stable sort and the scattered feature selection all run on the GPU, many views in flight.
 * template_ids[v] receives the new template_id or -1 (some level lacks candidates), bounding_boxes (nullable) the
 * cropTemplates boxes.  Result identical to n_views lm_add_template calls in order. */
int lm_add_templates_batch(lm_detector* det, const lm_image* sources, const lm_image* masks, int n_views, int n_sources,
                           const char* class_id, int32_t* template_ids, lm_rect* bounding_boxes /*nullable*/);
/* Render + addTemplate for n_views views without leaving the device: the trainer's loop.  mask_rects (nullable) receives
 * the render rectangles (`rects` at renderer.cpp:318), centre_depth_mm (nullable) the rendered depth at the image centre
 * (what `distance` is computed from at renderer.cpp:271). */
int lm_train_views(lm_detector* det, const lm_mesh* mesh, const lm_camera* cam, const double* T, const double* up,
                   int n_views, const char* class_id, int32_t* template_ids, lm_rect* bounding_boxes /*nullable*/,
                   lm_rect* mask_rects /*nullable*/, uint16_t* centre_depth_mm /*nullable*/);

/* The pose table the trainer writes next to templates.yml and the detection nodes read back:
 *   writeLinemodTemplateParams   src/renderer.cpp:72-123      "Template i": {ID, R, T, K, D, Ori_dist, Rect} + renderer_* keys
 *   readLinemodTemplateParams    src/rgbdDetector.cpp:1681-1749
 * Same cv::FileStorage YAML 1.0 layout (the shipped config/data/..._renderer_params.yml parses; files written here load
 * with cv::FileStorage).  Host only. */
typedef struct {
  double R[9];      /* Rs_: row-major, == lm_view_pose's R */
  double T[3];      /* Ts_: minus the camera position in the object frame (renderer_iterator.T()) */
  float K[9];       /* Ks_: [fx 0 cols/2; 0 fy rows/2; 0 0 1] (renderer.cpp:273) */
  double D;         /* distances_: D_obj - depth(image centre) / 1000 */
  double ori_dist;  /* Origin_dists_: D_obj, the view's radius */
  lm_rect rect;     /* rects */
} lm_template_pose;
typedef struct {
  int32_t n_points, angle_step;
  double radius_min, radius_max, radius_step;
  int32_t width, height;
  double fx, fy, near_, far_;
} lm_renderer_params;
int lm_write_renderer_params(const char* path, const lm_template_pose* poses, size_t n, const lm_renderer_params* params);
int lm_read_renderer_params(const char* path, lm_template_pose** out_poses, size_t* out_n, lm_renderer_params* params);
void lm_free_poses(lm_template_pose* poses);

/* ------------------------------------------------------------------------------------------------ hypothesis checks */
/* rgbdDetector::depth_diff as driven by depth_normal_diff_calc (src/rgbdDetector.cpp:147-283): for each
 * hypothesis i the template view (T, up) is rendered depth-only, cropped to its mask's bounding box, laid over the scene
 * depth at (x[i], y[i]) and the mean |template - scene| (metres) over pixels valid in both is returned in out[i]
 * (NaN when no pixel is valid, like the reference's 0/0).  LM_E_INVALID when a crop leaves the scene image. */
int lm_depth_diff_batch(lm_detector* det, const lm_image* scene_depth, const lm_mesh* mesh, const lm_camera* cam,
                        const double* T, const double* up, const int32_t* x, const int32_t* y, int n, double* out);

/* ------------------------------------------------------------------------------------------------ matching */
/* Detector::match(sources, threshold, matches, class_ids, quantized_images, masks)   src/rgbdDetector.cpp:33
 *   class_ids/n_ids   : empty = all classes in std::map order
 *   masks/n_masks     : 0 or n_sources images (LM_8UC1)
 *   quantized_out     : nullable; levels*M caller-allocated LM_8UC1 images (index l*M+m) of the level's size
 *   out_matches/out_n : library-allocated, release with lm_free_matches
 * Blocking; sorted and de-duplicated exactly like the reference (std::sort + std::unique on Match). */
int lm_match(lm_detector* det, const lm_image* sources, int n_sources, float threshold, const char* const* class_ids,
             int n_ids, const lm_image* masks, int n_masks, lm_image_out* quantized_out, lm_match_rec** out_matches,
             size_t* out_n);
/* Several (class list, threshold) queries answered from ONE front end of the frame -- what the reference's service
 * does with two detectors on the same image (object "memoryChip2" at 92, "cpu_binary" at 94,
 * /root/reference/launch/start_object_detection.launch:8,19), without quantising the frame twice.
 * out_offsets receives n_queries+1 prefix offsets into out_matches; each query's slice is sorted / de-duplicated
 * exactly like a separate lm_match call with that class list and threshold. */
typedef struct {
  float threshold;
  const char* const* class_ids; /* n_ids == 0: all classes */
  int n_ids;
} lm_query;
int lm_match_multi(lm_detector* det, const lm_image* sources, int n_sources, const lm_query* queries, int n_queries,
                   const lm_image* masks, int n_masks, lm_image_out* quantized_out, lm_match_rec** out_matches,
                   size_t* out_offsets);
/* The same over a batch of frames (sources[f*n_sources + m]).  Frames are processed in chunks of "batch_frames" (option,
 * default 8): every kernel launch of the path covers a whole chunk, and chunks are pipelined over "batch_lanes" workspace
 * lanes so that the host->device copies, the kernels and the result download of consecutive chunks overlap.
 * out_offsets receives n_frames+1 prefix offsets into out_matches. */
int lm_match_batch(lm_detector* det, const lm_image* sources, int n_frames, int n_sources, float threshold,
                   const char* const* class_ids, int n_ids, lm_match_rec** out_matches, size_t* out_offsets);
/* Batch + multi-query: every frame answers every query from one front end (a video stream watched by the reference's
 * two-object service).  out_offsets receives n_frames * n_queries + 1 prefix offsets, frame-major. */
int lm_match_batch_multi(lm_detector* det, const lm_image* sources, int n_frames, int n_sources, const lm_query* queries,
                         int n_queries, lm_match_rec** out_matches, size_t* out_offsets);
void lm_free_matches(lm_match_rec* matches);

/* A continuous stream of HOST frames: the same chunked pipeline as lm_match_batch_multi, kept alive between calls, so that
 * the device does not drain and refill at every call boundary (a 64-frame lm_match_batch_multi call spends about a fifth of
 * its time filling and draining the pipeline).  The reference's service calls Detector::match once per camera frame
 * (src/rgbdDetector.cpp:31-34); this is the same call for a caller that has the next frames already.
 *   lm_stream_open   fixes the queries (class ids are copied) and takes over the handle's workspace lanes: other matching
 *                    calls on the handle return LM_E_STATE until lm_stream_close.
 *   lm_stream_push   enqueues n_frames frames (sources[f*n_sources + m]) in chunks of "stream_frames" (default 16); returns once the
 *                    copies and kernels are enqueued, blocking only while all "batch_lanes" lanes are busy.  Pinned source
 *                    buffers are read asynchronously: they must stay unchanged until lm_stream_pop has returned their frames
 *                    (pageable ones are staged during the call).
 *   lm_stream_pop    hands out finished frames in push order: at most max_frames of them; wait_all = 0 returns what is ready
 *                    without blocking, wait_all = 1 first waits for everything pushed.  out_offsets (max_frames * n_queries
 *                    + 1 entries) and *out_matches as in lm_match_batch_multi (release with lm_free_matches); the lists are
 *                    identical to lm_match_multi's of the same frames.
 *   lm_stream_in_flight   frames pushed and not yet popped.
 * A stream must be closed before its detector is destroyed. */
typedef struct lm_stream lm_stream;
int lm_stream_open(lm_detector* det, const lm_query* queries, int n_queries, lm_stream** out);
int lm_stream_push(lm_stream* stream, const lm_image* sources, int n_frames, int n_sources);
int lm_stream_pop(lm_stream* stream, int wait_all, int max_frames, lm_match_rec** out_matches, size_t* out_offsets,
                  int* n_frames_out);
int lm_stream_in_flight(const lm_stream* stream);
void lm_stream_close(lm_stream* stream);

/* Device-resident variant for multi-GPU sharding and for callers that already hold the frame in HBM:
 * sources are DEVICE pointers (tightly packed rows), work is enqueued on `stream` (a cudaStream_t) and nothing is
 * synchronised.  The un-ordered survivor records stay in device memory: *d_records points at a header
 * {uint32 count, uint32 capacity, uint32 overflow, uint32 pad} followed by `capacity` lm_raw_match records.
 * lm_finalize_raw() (host) turns downloaded raw records (possibly concatenated from several template shards) into
 * the reference's sorted / de-duplicated match list. */
typedef struct {
  uint32_t order_key; /* bits 0-27: position of the template in the query's (class, template_id) iteration order;
                         bits 28-31: index of the query within the request (0 for single-query calls) */
  uint32_t coarse_pos; /* raster index r*W+c of the coarse candidate at the lowest pyramid level */
  int32_t x, y;        /* position after local refinement */
  uint32_t score;      /* raw integer similarity of the last evaluated level */
  uint32_t nf;         /* number of features behind `score` (similarity = score*100/(4*nf) [+0.5 if levels==1]) */
  int32_t template_id, class_index;
} lm_raw_match;

int lm_match_device(lm_detector* det, const void* const* d_sources, int n_sources, int rows, int cols, float threshold,
                    const char* const* class_ids, int n_ids, void* stream, const void** d_records,
                    size_t* record_bytes_capacity);
/* Multi-query form: one record block for the whole request; a record's query index is order_key >> 28. */
int lm_match_device_multi(lm_detector* det, const void* const* d_sources, int n_sources, int rows, int cols,
                          const lm_query* queries, int n_queries, void* stream, const void** d_records,
                          size_t* record_bytes_capacity);
/* The same with an explicit workspace lane (0..7): a handle owns eight independent workspaces + result blocks, so several
 * requests can be in flight on streams of the caller. */
int lm_match_device_multi_lane(lm_detector* det, int lane, const void* const* d_sources, int n_sources, int rows, int cols,
                               const lm_query* queries, int n_queries, void* stream, const void** d_records,
                               size_t* record_bytes_capacity);
/* The record blocks of a lane's chunk are one device allocation: block of frame f = *base + f * *frame_stride (each a
 * header + records as above; the bytes between blocks are padding and 16 bytes of kernel statistics).  A sharded caller
 * can exchange the survivors of a whole chunk with one collective over the region and no staging copies.  The region
 * moves only when the record capacity has to grow ("device_out_cap" option: records per block on the device-resident
 * path, default 2048) or a larger chunk is requested. */
int lm_device_result_region(lm_detector* det, int lane, const void** base, size_t* frame_stride, int* n_frames);
/* Enqueues on `stream` a device-to-device copy of the first `bytes` of the lane's first record block (header + leading
 * records) to d_dst: how a sharded caller parks the survivors of a frame in its send buffer without leaving the stream. */
int lm_copy_result_block(lm_detector* det, int lane, void* d_dst, size_t bytes, void* stream);
/* A run of device-resident frames in one call (sources d_sources[f * n_sources + m]): the frames are cut into chunks of
 * "batch_frames"; chunk c -- ONE launch set for all its frames -- is enqueued on lane c % n_streams and stream
 * streams[c % n_streams]; when d_stage is given, the head of frame f's record block (stage_slot_bytes: header + leading
 * records) is copied to d_stage + f * stage_slot_bytes on the same stream -- the send buffer of the sharded exchange (one
 * collective per run of frames).  Nothing is synchronised, nothing is copied: the kernels read the caller's buffers
 * through a device-resident frame table. */
int lm_match_device_stream(lm_detector* det, const void* const* d_sources, int n_frames, int n_sources, int rows, int cols,
                           const lm_query* queries, int n_queries, void* const* streams, int n_streams, void* d_stage,
                           size_t stage_slot_bytes);
/* Host -> device copies of n images into caller-owned device buffers (tightly packed rows), enqueued on `stream`: how a
 * sharded caller fills the chunk buffer it broadcasts without one binding call per image.  Page-locked sources are copied
 * asynchronously (they must stay valid until the stream reaches the copy); pageable ones synchronously. */
int lm_upload_images(lm_detector* det, const lm_image* images, int n, void* const* d_dst, void* stream);
int lm_finalize_raw(const lm_detector* det, const lm_raw_match* raw, size_t n_raw, lm_match_rec** out_matches,
                    size_t* out_n);
/* The same for a whole exchange buffer of a streamed run: blocks + r * rank_stride + f * block_bytes is rank r's staged
 * block of frame f (header + leading records, as lm_match_device_stream parks them and an all-gather concatenates them).
 * For every frame and query the union of the ranks' records is finalised like lm_finalize_raw; out_offsets receives
 * n_frames * n_queries + 1 prefix offsets (frame-major).  frame_status[f]: 0 = finalised, 1 = some rank has more than
 * capacity_records survivors for this frame (its slices are empty: redo the frame with a larger exchange), 2 = a rank's
 * candidate list overflowed on the device. */
int lm_finalize_gathered(const lm_detector* det, const void* blocks, int world, int n_frames, size_t block_bytes,
                         size_t rank_stride, uint32_t capacity_records, int n_queries, lm_match_rec** out_matches,
                         size_t* out_offsets, uint8_t* frame_status);
/* Template sharding (north-star multi-GPU layout): keep only templates whose canonical order index i satisfies
 * i % world == rank on this handle; template_id / class_index / order_key stay global. */
int lm_set_shard(lm_detector* det, int rank, int world);

/* ------------------------------------------------------------------------------------------------ several GPUs, one caller */
/* The multi-GPU handle for the reference's kind of caller -- ONE process making one call
 * (rgbdDetector::linemod_detection, src/rgbdDetector.cpp:31-34): the prototype's model (templates, modalities, T, tables,
 * options) is cloned onto every listed device and each device is driven by its own worker thread.
 *   LM_GROUP_FRAMES     every device holds all templates; the frames of a batch are dealt out in launch sets ("batch_frames"
 *                       frames, round robin), each device copies its frames in over its own PCIe link and returns finished
 *                       lists: no exchange between devices, the end-to-end frame rate scales with the device count.
 *   LM_GROUP_TEMPLATES  the north-star layout: templates sharded by canonical index, every device sees every frame, the
 *                       shards' survivors are merged before the reference's sort + unique.  Single-frame latency, or template
 *                       sets that outgrow one device.
 * Results are identical to lm_match_batch_multi on the prototype in both modes.  (One process per GPU with an NCCL
 * exchange instead: linemod_pose_estimation_b200/sharding.py over lm_set_shard / lm_match_device_stream / lm_finalize_gathered.) */
typedef struct lm_group lm_group;
#define LM_GROUP_FRAMES 0
#define LM_GROUP_TEMPLATES 1
int lm_group_create(const lm_detector* prototype, const int* devices, int n_devices, int mode, lm_group** out);
/* Both at once (a 2-D grid): the devices form n_devices / template_shards sets of `template_shards` template shards
 * (device i = shard i % S of set i / S); launch sets of frames are dealt out to the sets, the devices of a set see the set's
 * frames and their survivors are merged.  template_shards = 1 is LM_GROUP_FRAMES, = n_devices is LM_GROUP_TEMPLATES. */
int lm_group_create_grid(const lm_detector* prototype, const int* devices, int n_devices, int template_shards, lm_group** out);
void lm_group_destroy(lm_group* group);
int lm_group_size(const lm_group* group);
int lm_group_mode(const lm_group* group);
lm_detector* lm_group_member(lm_group* group, int i);              /* the handle bound to devices[i] (introspection, options) */
int lm_group_set_option(lm_group* group, const char* key, int value); /* lm_set_option on every member */
/* lm_match_batch_multi / lm_match across the group's devices; same outputs, released with lm_free_matches */
int lm_group_match_batch_multi(lm_group* group, const lm_image* sources, int n_frames, int n_sources, const lm_query* queries,
                               int n_queries, lm_match_rec** out_matches, size_t* out_offsets);
int lm_group_match(lm_group* group, const lm_image* sources, int n_sources, float threshold, const char* const* class_ids,
                   int n_ids, lm_match_rec** out_matches, size_t* out_n);

/* ------------------------------------------------------------------------------------------------ match clustering */
/* The stage right behind Detector::match in the reference's nodes:
 *   rgbdDetector::rcd_voting            src/rgbdDetector.cpp:36-70    bin matches by (y / step, x / step, depth bin of the template)
 *   rgbdDetector::cluster_filter        src/rgbdDetector.cpp:72-84    drop bins with <= cluster_threshold matches
 *   rgbdDetector::similarity_score_calc src/rgbdDetector.cpp:133-145  cluster score = mean similarity
 *   rgbdDetector::nonMaximaSuppressionUsingIOU + computeIoU   src/rgbdDetector.cpp:462-574   mean rectangle per cluster,
 *                                       std::sort by score (descending), greedy suppression at IoU > iou_threshold
 * as driven at src/linemod_ensenso_detect_3_mult_detect.cpp:352-424 (thresh = 2, IoU 0.4).  Host code: a frame yields a few
 * dozen matches.  (cluster_filter erases from the std::map while iterating over it -- undefined behaviour upstream; the
 * evident intent, "erase every bin with size <= thresh", is what is implemented.) */
typedef struct {
  int32_t vote_row_col_step;     /* clustering_step_: bin size in pixels (rows and columns) */
  double renderer_radius_min;    /* depth of bin 0 (m) */
  double renderer_radius_step;   /* depth bin size (m); converted to float like the reference does */
  int32_t cluster_threshold;     /* bins with <= this many matches are dropped (reference: 2) */
  double iou_threshold;          /* reference: 0.4 */
} lm_cluster_params;
typedef struct {
  int32_t index[3];              /* (row bin, column bin, depth bin) */
  double score;                  /* mean similarity of the cluster's matches */
  lm_rect rect;                  /* mean x, y of the matches; mean width, height of their templates' rectangles */
  uint32_t first, count;         /* the cluster's matches: match_index[first .. first + count) */
} lm_cluster;
/* obj_origin_dists / rects are indexed by template_id (the reference's renderer_params.yml tables).  Outputs are
 * library-allocated (release both with lm_free_clusters): the surviving clusters in the reference's order and, grouped
 * per cluster, indices into `matches` in the order the reference's std::vector<Match> holds them. */
int lm_cluster_matches(const lm_match_rec* matches, size_t n_matches, const double* obj_origin_dists, const lm_rect* rects,
                       size_t n_templates, const lm_cluster_params* params, lm_cluster** out_clusters, size_t* out_n,
                       uint32_t** out_match_index);
void lm_free_clusters(lm_cluster* clusters, uint32_t* match_index);

/* ------------------------------------------------------------------------------------------------ data tables */
/* SIMILARITY_LUT[256] ([OCV] linemod.cpp) and NORMAL_LUT[20][20][20] ([OCV] normal_lut.i) are data, not code.  The first
 * is the literal upstream table (tests/golden/similarity_lut_ocv.txt); the second is not recoverable without OpenCV's
 * sources and a stand-in is shipped (DESIGN.md section 2 "LUTs"), so both are injectable.  Similarity entries must be <= 4. */
int lm_set_similarity_lut(lm_detector* det, const uint8_t lut[256]);
int lm_get_similarity_lut(const lm_detector* det, uint8_t lut[256]);
int lm_set_normal_lut(lm_detector* det, const uint8_t lut[8000]);
int lm_get_normal_lut(const lm_detector* det, uint8_t lut[8000]);
/* The same from OpenCV's own text file (modules/objdetect/src/normal_lut.i, a brace-initialised unsigned char
 * [20][20][20]): every integer after the first '{' is an entry; exactly 8000 entries <= 255 are required (LM_E_IO). */
int lm_load_normal_lut_file(lm_detector* det, const char* path);

/* ------------------------------------------------------------------------------------------------ parity taps */
#define LM_STAGE_QUANTIZED 0 /* u8 [rows][cols], after mask                      (QuantizedPyramid::quantize) */
#define LM_STAGE_SPREAD 1    /* u8 [rows][cols]                                  (spread) */
#define LM_STAGE_RESPONSE 2  /* u8 [8][rows][cols]                               (computeResponseMaps) */
#define LM_STAGE_LINEAR 3    /* u8 [8][plane_stride]                             (linearize, flat + zero tail) */
#define LM_STAGE_MAGNITUDE 4 /* f32 [rows][cols], ColorGradient only             (quantizedOrientations) */
#define LM_STAGE_QUANT_RAW 5 /* u8 [rows][cols], before mask */
#define LM_STAGE_LINEAR_PACKED 6 /* u8 [8][plane_stride / 2]: the coarsest level's LM_STAGE_LINEAR packed two positions
                                    per byte (position p = nibble p), the layout the matching kernels read */
/* Copies a stage of the LAST lm_match / lm_build_front call to host memory. dst NULL = size query. Returns bytes. */
long lm_debug_fetch(lm_detector* det, int stage, int level, int modality, void* dst);
/* Front end only (quantise -> spread -> response -> linearize) without matching. */
int lm_build_front(lm_detector* det, const lm_image* sources, int n_sources, const lm_image* masks, int n_masks);
/* rows, cols, T, W, H of a level and the byte stride between orientation planes in LM_STAGE_LINEAR. */
int lm_level_geometry(lm_detector* det, int level, int32_t out[5], size_t* plane_stride);
/* Coarse u16 similarity map [H][W] of one template against the last front end (matchClass before thresholding). */
int lm_debug_coarse_map(lm_detector* det, const char* class_id, int template_id, uint16_t* dst);
/* Raw (unsorted, order-restored) matches of the last lm_match call, i.e. the list handed to std::sort. */
long lm_debug_presort(lm_detector* det, lm_match_rec* dst /*nullable*/);

/* ------------------------------------------------------------------------------------------------ measurement */
/* Device time (ms, CUDA events on the detector's stream) of the stages of the LAST lm_match / lm_match_multi call made with
 * the "timing" option on (plain launches with events between the stages instead of the lane's CUDA graph):
 * [0] H2D, [1] front end, [2] coarse similarity, [3] local refinement, [4] D2H; and the kernel launches of the last call. */
int lm_last_timings(const lm_detector* det, float ms[5], int* kernel_launches);
/* Algorithmic bytes of the last launch set on lane 0 -- the last lm_match call, or the last chunk of
 * an lm_match_batch* call that lane 0 processed: [0] B_front per frame, [1] B_coarse [2] B_refine [3] B_out of all its frames,
 * [4] coarse candidates, [5] template*position evals, [6] the part of B_coarse the coarse kernel actually gathered (exact
 * early termination skips features of tiles in which no position can reach the threshold any more), [7] frames in the set. */
int lm_last_work(const lm_detector* det, uint64_t out[8]);
/* Tuning switches: "batch_frames" (frames per chunk = per launch set on the batched paths, 1..32, default 8),
 * "batch_lanes" (chunks in flight in lm_match_batch*, default 4), "finalize_threads" (host threads that order the match
 * lists of batched calls while the calling thread feeds the device, default 4; 0 = on the calling thread), "prune" (exact early termination: bit 0 = in the coarse
 * kernel, bit 1 = of hopeless candidates in the refinement kernel; default 3; results do not depend on it), "mod_order" (order in which the coarse kernel sums the modalities: 0 = template order, 1 = reversed,
 * 2 = chosen per frame from the front end's spread-bit counters, default; results do not depend on it), "graphs" (replay
 * a recorded CUDA graph per chunk, default 1), "timing" (per-stage events for lm_last_timings, default 0), "debug_taps",
 * "stream_frames" (frames per chunk of an lm_stream opened afterwards, 1..32, default 16), "refine_tiled" (refinement levels
 * keep their nibble planes column-blocked, default 1), "coarse_narrow" (requests whose tiles have at most 63 features use the
 * u8-only coarse kernel: 1 default, 0 the general kernel, 2 the u8-only body at two CTAs per SM; process-wide),
 * "coarse_share" (coarse tail passes of at most 128 positions are scored for eight frames per warp, default 1), "dn_count"
 * (DepthNormal's medianBlur(5) by counting when the NORMAL_LUT is one-hot, default 1; 0 keeps the median network) -- A/B
 * switches, results do not depend on them --
 * "coarse_grid_limit", "device_out_cap" (records per frame block on the device-resident paths, default 2048),
 * "cand_per_frame" (coarse candidates a chunk may produce per frame on the device-resident paths, default 65536; the
 * host paths grow both by themselves). */
int lm_set_option(lm_detector* det, const char* key, int value);

#ifdef __cplusplus
}
#endif
#endif /* LINEMOD_B200_H_ */
