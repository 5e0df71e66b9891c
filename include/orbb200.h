/* orbb200 -- C ABI of the B200-native ORB front-end (extraction + Hamming matching).
 *
 * Drop-in boundary for the data-parallel hot path of donglinb/ORB-SLAM-BIRDVIEW.  The reference has no
 * FFI layer: the boundary is two C++ classes, ORB_SLAM2::ORBextractor (include/ORBextractor.h:44-111) and
 * ORB_SLAM2::ORBmatcher (include/ORBmatcher.h:38-120).  The host shims in
 * orb-slam-birdview_b200/cpp/ keep those class signatures and call the functions below; each entry cites
 * the reference code it replaces.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; return 0 on success, a negative
 * orbb200_status otherwise (never throws or aborts); all buffers are caller-owned.  "host" entry points
 * take host pointers and are synchronous; "_device" entry points take device pointers, enqueue on the
 * context's stream and return immediately (orbb200_sync waits).  A context is used by one host thread at
 * a time (the reference uses one ORBextractor per camera/thread: src/Tracking.cc:121-127,
 * src/Frame.cc:124-127); any number of contexts may exist per process and per GPU.
 * There is no CPU fallback: without a CUDA device every call fails with ORBB200_ERR_CUDA.
 */
#ifndef ORBB200_H
#define ORBB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    ORBB200_OK = 0,
    ORBB200_ERR_CUDA = -1,        /* CUDA runtime error or no device; see orbb200_last_error */
    ORBB200_ERR_ARG = -2,         /* bad argument (null pointer, size over the context's limits, ...) */
    ORBB200_ERR_CAPACITY = -3,    /* caller buffer too small for the result */
    ORBB200_ERR_UNSUPPORTED = -4  /* geometry outside what the reference defines */
} orbb200_status;

/* Binary layout of cv::KeyPoint (28 bytes): pt.x, pt.y, size, angle, response, octave, class_id */
typedef struct { float x, y, size, angle, response; int32_t octave, class_id; } orbb200_kp_t;

typedef struct orbb200_ctx orbb200_ctx;
typedef struct orbb200_frame orbb200_frame;

/* ---- context ------------------------------------------------------------------------------------
 * Replaces ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
 * (src/ORBextractor.cc:410-470).  max_w/max_h/max_batch size the device buffers (HBM resident pyramid,
 * blurred pyramid, candidate and keypoint pools for max_batch images). */
int orbb200_create(orbb200_ctx** out, int device, int nfeatures, float scaleFactor, int nlevels,
                   int iniThFAST, int minThFAST, int max_w, int max_h, int max_batch);
void orbb200_destroy(orbb200_ctx* ctx);
const char* orbb200_last_error(const orbb200_ctx* ctx);   /* ctx may be NULL: last create() error of the calling thread */
int orbb200_sync(orbb200_ctx* ctx);
/* cudaStream_t of the context, as void* (for callers that enqueue their own work around ours) */
void* orbb200_stream(orbb200_ctx* ctx);

/* GetLevels / GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares /
 * GetInverseScaleSigmaSquares (include/ORBextractor.h:63-82); which: 0..3 in that order */
int orbb200_get_levels(const orbb200_ctx* ctx);
int orbb200_get_scale_table(const orbb200_ctx* ctx, int which, float* out /*[nlevels]*/);
int orbb200_get_features_per_level(const orbb200_ctx* ctx, int32_t* out /*[nlevels]*/);
/* maximum keypoints one image can produce with this context (sum over levels of quota + slack) */
int orbb200_max_keypoints(const orbb200_ctx* ctx);

/* ---- extraction ---------------------------------------------------------------------------------
 * Replaces ORBextractor::operator()(image, mask, keypoints, descriptors) (src/ORBextractor.cc:1043-1105;
 * the mask is ignored by the reference too).  img: CV_8UC1 rows of `stride` bytes.  Writes *n_out
 * keypoints (reference order: level 0..L-1, each level in octree list order) and n_out*32 descriptor bytes. */
int orbb200_extract(orbb200_ctx* ctx, const uint8_t* img, int w, int h, size_t stride,
                    orbb200_kp_t* kps, uint8_t* desc, int cap, int* n_out);
/* n images of one shape in one pass (frame batches, stereo pairs).  Results for image i start at
 * kps[i*cap_per_img] / desc[i*cap_per_img*32]; n_out[i] keypoints each. */
int orbb200_extract_batch(orbb200_ctx* ctx, const uint8_t* const* imgs, int n, int w, int h, size_t stride,
                          orbb200_kp_t* kps, uint8_t* desc, int cap_per_img, int* n_out);
/* Device-resident variant: d_imgs holds n images, image i at d_imgs + i*img_bytes, rows of `stride` bytes.
 * Results stay in the context's device pools (orbb200_results_device) for the matchers. Asynchronous. */
int orbb200_extract_device(orbb200_ctx* ctx, const uint8_t* d_imgs, size_t img_bytes, int n, int w, int h, size_t stride);
/* Device pointers of the last extraction: kps [max_batch][cap] , desc [max_batch][cap][32], counts [max_batch] */
int orbb200_results_device(orbb200_ctx* ctx, const orbb200_kp_t** d_kps, const uint8_t** d_desc,
                           const int32_t** d_counts, int* cap_per_img);
/* Copy the results of the last extraction (either variant) to host buffers. */
int orbb200_download_results(orbb200_ctx* ctx, int n, orbb200_kp_t* kps, uint8_t* desc, int cap_per_img, int* n_out);
/* Lazy download of one pyramid level of image `img_index` of the last extraction: what the reference keeps
 * in the public member mvImagePyramid (include/ORBextractor.h:85; read by Frame::ComputeStereoMatches,
 * src/Frame.cc:669-776).  blurred!=0 returns the GaussianBlur'ed copy used for the descriptors. */
int orbb200_pyramid_level(orbb200_ctx* ctx, int img_index, int level, int blurred,
                          uint8_t* dst, size_t dst_stride, int* w, int* h);
/* The same for all levels at once: ONE device-to-host copy of the image's pyramid block into a pinned mirror owned by the
 * context, and per level the address of pixel (0,0) inside it with its row pitch (rows keep the device layout: the
 * reflect-101 border the reference's copyMakeBorder produces lies around every level, exactly as around the ROIs the reference
 * keeps in mvImagePyramid).  The pointers stay valid until the next call of this function on the context.  This is what the
 * ORBextractor shim fills mvImagePyramid with (cv::Mat headers over the mirror, no per-level copies). */
int orbb200_pyramid_mirror(orbb200_ctx* ctx, int img_index, int blurred, const uint8_t** level_ptr /*[nlevels]*/,
                           size_t* level_pitch /*[nlevels]*/, int* level_w /*[nlevels]*/, int* level_h /*[nlevels]*/);
/* mvImagePyramid without a copy: after orbb200_set_pyramid_mirror(ctx, 1), an orbb200_extract of one or two staged images also stores
 * image 0's levels into the pinned mirror from the kernels that produce them, and the next orbb200_pyramid_mirror(ctx, 0, 0, ...) returns
 * the level pointers without copying or synchronising (include/ORBextractor.h:85, read by Frame::ComputeStereoMatches). */
int orbb200_set_pyramid_mirror(orbb200_ctx* ctx, int enable);
/* Debug/inspection: FAST candidates of one level as packed (x,y,response) int32 triples in region
 * coordinates, unordered.  Returns the count (or a negative status). */
int orbb200_level_candidates(orbb200_ctx* ctx, int img_index, int level, int32_t* xyr, int cap);

/* ---- brute-force Hamming (C4) -------------------------------------------------------------------
 * ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1647-1663) over all nq x nm pairs; per query the best
 * index (first minimum in index order), best distance and second-best distance (256 when absent). */
int orbb200_hamming_knn2(orbb200_ctx* ctx, const uint8_t* q, int nq, const uint8_t* m, int nm,
                         int32_t* best_idx, int32_t* best_d, int32_t* second_d);
int orbb200_hamming_knn2_device(orbb200_ctx* ctx, const uint8_t* d_q, int nq, const uint8_t* d_m, int nm,
                                int32_t* d_best_idx, int32_t* d_best_d, int32_t* d_second_d);

/* ---- representative descriptor of a landmark -----------------------------------------------------
 * The selection step of MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:272-301) and
 * MapPointBird::ComputeDistinctiveDescriptors (src/MapPointBird.cc:117-146), batched over landmarks: group g is the
 * descriptors desc[group_ptr[g] .. group_ptr[g+1]) of one landmark's observations (in the order the reference
 * pushes them: std::map<KeyFrame*,size_t> iteration order, for MapPointBird preceded by its own descriptor when it
 * has no reference keyframe).  Per group: all pairwise DescriptorDistance, per row the element [0.5*(N-1)] of the
 * sorted row (self distance 0 included), best_idx = FIRST row with the smallest median, best_median = that value;
 * -1 / -1 for an empty group (the reference returns without touching mDescriptor). */
int orbb200_distinctive_descriptors(orbb200_ctx* ctx, const uint8_t* desc, const int32_t* group_ptr, int n_groups,
                                    int32_t* best_idx, int32_t* best_median);

/* ---- frames: keypoints + descriptors + 64x48 lookup grid on the device ---------------------------
 * Replaces Frame::AssignFeaturesToGrid / PosInGrid[Birdview] (src/Frame.cc:378-412,549-559,879-889) and is
 * what Frame::GetFeaturesInArea[Birdview] (src/Frame.cc:494-547,891-944) scans.  Front camera:
 * (min_x,min_y,inv_w,inv_h) = (mnMinX,mnMinY,mfGridElementWidthInv,mfGridElementHeightInv); birdview:
 * (0,0,mfGridElementWidthInvBirdview,mfGridElementHeightInvBirdview).  u_right may be NULL (all -1). */
int orbb200_frame_upload(orbb200_ctx* ctx, orbb200_frame** f, const orbb200_kp_t* kps, const uint8_t* desc,
                         const float* u_right, int n, float min_x, float min_y, float inv_w, float inv_h);
/* Frame over image `img_index` of the last extraction without leaving the device. */
int orbb200_frame_from_extract(orbb200_ctx* ctx, orbb200_frame** f, int img_index,
                               float min_x, float min_y, float inv_w, float inv_h);
void orbb200_frame_free(orbb200_frame* f);
/* GetFeaturesInArea for one query (testing aid; order = reference scan order).  Returns the count. */
int orbb200_frame_features_in_area(orbb200_ctx* ctx, const orbb200_frame* f, float x, float y, float r,
                                   int min_level, int max_level, int32_t* out, int cap);

/* ---- windowed searches ----------------------------------------------------------------------------
 * All take host arrays of nq flattened queries and write host results.  q_desc is [nq][32].
 * *_obs_pos / kp_blocked carry the loop-carried "keypoint already holds a map point with
 * Observations()>0" state of the reference loops (may be NULL: all 1 / all 0). */

/* ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:45-129).
 * out_query_of_kp[n]: query finally assigned to each keypoint (F.mvpMapPoints), -1 none. Returns nmatches
 * through *nmatches. */
int orbb200_search_by_projection(orbb200_ctx* ctx, const orbb200_frame* F, int nq,
                                 const uint8_t* q_valid, const float* q_u, const float* q_v, const float* q_uR,
                                 const int32_t* q_level, const float* q_viewcos, const uint8_t* q_desc,
                                 const uint8_t* q_obs_pos, const uint8_t* kp_blocked, float th, float nnratio,
                                 int32_t* out_best_idx, int32_t* out_best_dist, int32_t* out_query_of_kp,
                                 int* nmatches);

/* ORBmatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, bMono) (src/ORBmatcher.cc:1328-1470);
 * the caller projects Last's map points (u, v, 1/z).  mode 0: levels [oct-1,oct+1]; 1: forward (>=oct);
 * 2: backward ([0,oct]).  check_ori: rotation-histogram filter (:1431-1467). */
int orbb200_search_by_projection_frame(orbb200_ctx* ctx, const orbb200_frame* Cur, int nq,
                                       const uint8_t* q_valid, const float* q_u, const float* q_v, const float* q_invz,
                                       const int32_t* q_octave, const float* q_angle, const uint8_t* q_desc,
                                       const uint8_t* q_obs_pos, const uint8_t* kp_blocked,
                                       float th, float mbf, int mode, int check_ori,
                                       int32_t* out_query_of_kp, int* nmatches);

/* ORBmatcher::BirdviewMatch(F1,F2,vnMatches12,vPrevMatched,windowSize) (src/ORBmatcher.cc:1667-1786) when
 * prev_xy != NULL (octave-0 queries, window around prev_xy[i], prev_xy updated in place), and
 * ORBmatcher::BirdviewMatch(const F1,const F2,vnMatches12,windowSize) (:1788-1899) when prev_xy == NULL.
 * F1 is given by its keypoints/descriptors (host), F2 as a device frame. */
int orbb200_birdview_match(orbb200_ctx* ctx, const orbb200_kp_t* kps1, const uint8_t* desc1, int n1,
                           const orbb200_frame* F2, float* prev_xy, int window_size, float nnratio, int check_ori,
                           int32_t* matches12, int* nmatches);

/* ORBmatcher::SearchByMatchBird(KeyFrame*, Frame&, vector<MapPointBird*>&, r) (src/ORBmatcher.cc:2000-2114) */
int orbb200_search_by_match_bird_kf(orbb200_ctx* ctx, const orbb200_kp_t* kf_kps, const uint8_t* has_mp,
                                    const uint8_t* mp_desc, int nk, const orbb200_frame* F, float r, float nnratio,
                                    int check_ori, int32_t* out_mp_of_kp, int* nmatches);

/* ORBmatcher::SearchByProjectionBird(Frame&, const vector<MapPointBird*>&, r) (src/ORBmatcher.cc:1923-1998);
 * the caller projects the landmarks through Tbc*Tcw and sets q_valid. */
int orbb200_search_by_projection_bird(orbb200_ctx* ctx, const orbb200_frame* F, int nq, const uint8_t* q_valid,
                                      const float* q_x, const float* q_y, const uint8_t* q_desc,
                                      const uint8_t* q_obs_pos, const uint8_t* kp_blocked, float r, float nnratio,
                                      int32_t* out_query_of_kp, int* nmatches);

/* ORBmatcher::SearchForTriangulation(KF1,KF2,F12,pairs,bOnlyStereo) (src/ORBmatcher.cc:657-823).  Feature
 * vectors as CSR over ascending vocabulary node ids.  pairs: [n1][2] (idx1, idx2), *npairs written. */
int orbb200_search_for_triangulation(orbb200_ctx* ctx,
                                     const orbb200_kp_t* kps1, const uint8_t* desc1, const float* uR1, const uint8_t* has_mp1, int n1,
                                     const orbb200_kp_t* kps2, const uint8_t* desc2, const float* uR2, const uint8_t* has_mp2, int n2,
                                     const int32_t* fv1_node, const int32_t* fv1_ptr, const int32_t* fv1_idx, int nn1,
                                     const int32_t* fv2_node, const int32_t* fv2_ptr, const int32_t* fv2_idx, int nn2,
                                     const float* F12, float ex, float ey,
                                     const float* scale_factors2, const float* level_sigma2_2,
                                     int only_stereo, int check_ori, int32_t* pairs, int* npairs);

/* ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (src/ORBmatcher.cc:405-520):
 * octave-0 keypoints of F1, window around vbPrevMatched[i1] in F2's grid; prev_xy updated in place. */
int orbb200_search_for_initialization(orbb200_ctx* ctx, const orbb200_kp_t* kps1, const uint8_t* desc1, int n1,
                                      const orbb200_frame* F2, float* prev_xy, int window_size, float nnratio, int check_ori,
                                      int32_t* matches12, int* nmatches);

/* Generic "best descriptor in a window" search shared by the remaining projection searches.  The caller (the
 * ORBmatcher adapter) does the geometry on the host exactly as the reference does and passes, per query, the
 * window centre (q_x,q_y), radius q_r and level range [q_min_level,q_max_level] (reference conventions of
 * GetFeaturesInArea: -1 = unbounded).  flags:
 *   ORBB200_WB_BLOCK    an accepted match blocks its keypoint for later queries (loop-carried state)
 *   ORBB200_WB_URCHECK  skip a keypoint with uRight>0 when |q_aux - uRight| > radius
 *   ORBB200_WB_CHI2     Fuse's reprojection gate: (q_x,q_y,q_aux=ur) vs the keypoint, 5.99 / 7.8 * sigma2[level]
 *   ORBB200_WB_ORI      rotation-histogram filter on the accepted matches (q_angle)
 * Accept when best <= acc_th.  Covers:
 *   SearchByProjection(Frame&,KeyFrame*,set,th,ORBdist) (:1472-1599): r=th*s[pred], levels pred-1..pred+1, BLOCK|ORI, ORBdist
 *   SearchByProjection(KeyFrame*,Scw,points,vpMatched,th) (:290-403): levels pred-1..pred, BLOCK, TH_LOW
 *   Fuse(KeyFrame*,points,th) / Fuse(KeyFrame*,Scw,...) (:825-1100): levels pred-1..pred, CHI2 (first form), TH_LOW, independent
 *   SearchBySim3 (:1102-1326), each direction: levels pred-1..pred, TH_HIGH, independent, kp_blocked = vbAlreadyMatched */
#define ORBB200_WB_BLOCK 1
#define ORBB200_WB_URCHECK 2
#define ORBB200_WB_CHI2 4
#define ORBB200_WB_ORI 8
int orbb200_search_window_best(orbb200_ctx* ctx, const orbb200_frame* F, int nq, const uint8_t* q_valid,
                               const float* q_x, const float* q_y, const float* q_r, const int32_t* q_min_level,
                               const int32_t* q_max_level, const uint8_t* q_desc, const float* q_aux, const float* q_angle,
                               const uint8_t* q_obs_pos, const uint8_t* kp_blocked, const float* inv_level_sigma2,
                               int acc_th, int flags, int32_t* out_best_idx, int32_t* out_best_dist,
                               int32_t* out_query_of_kp, int* nmatches);

/* ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (src/ORBmatcher.cc:159-288) when kf_kf == 0 and
 * ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (:522-655) when kf_kf != 0.  valid1[i]: KF1 keypoint i
 * has a good MapPoint; valid2 (kf_kf only): same for KF2.  Feature vectors as CSR over ascending node ids.
 * out: kf_kf == 0 -> [n2] KF keypoint index whose MapPoint each F keypoint received (-1 none);
 *      kf_kf != 0 -> [n1] KF2 keypoint index matched to each KF1 keypoint (-1 none). */
int orbb200_search_by_bow(orbb200_ctx* ctx, const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1,
                          const orbb200_frame* F2, const uint8_t* valid2,
                          const int32_t* fv1_node, const int32_t* fv1_ptr, const int32_t* fv1_idx, int nn1,
                          const int32_t* fv2_node, const int32_t* fv2_ptr, const int32_t* fv2_idx, int nn2,
                          float nnratio, int check_ori, int kf_kf, int32_t* out, int* nmatches);

/* ---- stereo matching -------------------------------------------------------------------------------------
 * Frame::ComputeStereoMatches (src/Frame.cc:662-836) on the results of the last extraction: row-band descriptor
 * search, 11x11 SAD refinement on the pyramid level of the left keypoint, parabola sub-pixel fit, median outlier
 * cut.  Frame f uses images left0 + f*stride_imgs and right0 + f*stride_imgs of the extracted batch.  The device
 * keeps mvuRight / mvDepth per left image ([max_batch][cap]); mb = baseline, mbf = baseline*fx. */
int orbb200_stereo_matches_device(orbb200_ctx* ctx, int n_frames, int left0, int right0, int stride_imgs, float mb, float mbf);
int orbb200_stereo_results_device(orbb200_ctx* ctx, const float** d_uright, const float** d_depth, const int32_t** d_nkept);
/* One stereo pair, results copied to the host: u_right[cap], depth[cap] (first N entries meaningful). */
int orbb200_compute_stereo_matches(orbb200_ctx* ctx, int img_left, int img_right, float mb, float mbf,
                                   float* u_right, float* depth, int cap, int* n_matches);
/* Frame over the left image of the last extraction whose uRight comes from the stereo matching above. */
int orbb200_frame_from_extract_stereo(orbb200_ctx* ctx, orbb200_frame** f, int img_left,
                                      float min_x, float min_y, float inv_w, float inv_h);

/* ---- local map on the device: Frame::isInFrustum + SearchLocalPoints ---------------------------------------
 * Tracking::SearchLocalPoints (src/Tracking.cc:1610-1660) projects every local map point with
 * Frame::isInFrustum (src/Frame.cc:436-492; MapPoint::PredictScale src/MapPoint.cc:402-417) and then calls
 * ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th) (src/ORBmatcher.cc:45-129).  The map snapshot
 * (world position, mean viewing direction, mfMaxDistance / mfMinDistance, representative descriptor per point) stays
 * on the device between frames, so a frame only sends its pose. */
typedef struct orbb200_map orbb200_map;
typedef struct {
    float Rcw[9], tcw[3], Ow[3];          /* Frame::mRcw (row-major), mtcw, mOw */
    float fx, fy, cx, cy, mbf;
    float min_x, max_x, min_y, max_y;     /* Frame::mnMinX .. mnMaxY */
    float log_scale_factor;               /* Frame::mfLogScaleFactor */
    int32_t n_levels;                     /* Frame::mnScaleLevels */
} orbb200_camera_pose;
int orbb200_map_upload(orbb200_ctx* ctx, orbb200_map** map, int n, const float* pos /*[n][3]*/, const float* normal /*[n][3]*/,
                       const float* max_distance /*[n] mfMaxDistance*/, const float* min_distance /*[n] mfMinDistance*/,
                       const uint8_t* desc /*[n][32]*/);
void orbb200_map_free(orbb200_map* map);
/* Frame::isInFrustum for every map point with candidate[i] != 0 (NULL: all; the caller clears points with
 * mnLastFrameSeen == frame id or isBad(), src/Tracking.cc:1638-1641).  Outputs (host, [n], any may be NULL) are the
 * MapPoint::mbTrackInView / mTrackProjX / mTrackProjY / mTrackProjXR / mnTrackScaleLevel / mTrackViewCos fields.
 * *n_in_view = nToMatch. */
int orbb200_is_in_frustum(orbb200_ctx* ctx, const orbb200_map* map, const orbb200_camera_pose* pose, float viewing_cos_limit,
                          const uint8_t* candidate, uint8_t* out_in_view, float* out_u, float* out_v, float* out_uR,
                          int32_t* out_level, float* out_viewcos, int* n_in_view);
/* isInFrustum + SearchByProjection(F, local map, th) without the queries leaving the device.  obs_pos / kp_blocked /
 * outputs as in orbb200_search_by_projection (queries = map points); out_in_view etc. as in orbb200_is_in_frustum. */
int orbb200_search_local_points(orbb200_ctx* ctx, const orbb200_frame* F, const orbb200_map* map,
                                const orbb200_camera_pose* pose, float viewing_cos_limit, const uint8_t* candidate,
                                const uint8_t* obs_pos, const uint8_t* kp_blocked, float th, float nnratio,
                                uint8_t* out_in_view, float* out_u, float* out_v, float* out_uR, int32_t* out_level,
                                float* out_viewcos, int* n_in_view,
                                int32_t* out_best_idx, int32_t* out_best_dist, int32_t* out_query_of_kp, int* nmatches);

/* ---- birdview front-end: cv::ORB + cv::cornerSubPix ------------------------------------------------------------
 * What Frame::Frame runs on the birdview image (src/Frame.cc:328-342):
 *     cv::Ptr<cv::ORB> e = cv::ORB::create(2000);  e->detect(img, kps, mask);
 *     cv::cornerSubPix(img, pts, Size(5,5), Size(-1,-1), TermCriteria(EPS + MAX_ITER, 40, 0.001));  e->compute(img, kps, desc);
 * cv::ORB defaults: scaleFactor 1.2f, 8 levels, edgeThreshold 31, HARRIS_SCORE, patchSize 31, fastThreshold 20, WTA_K 2.
 * Results follow OpenCV 4.x's own C++ code bit for bit (cv2 4.13 with cv2.setUseOptimized(False)), including the order
 * in which KeyPointsFilter::retainBest leaves the keypoints.  mask may be NULL; rows of `stride` / `mask_stride` bytes. */
/* Upper bound of the keypoints one image can return (size your buffers with it). */
int orbb200_bird_max_keypoints(orbb200_ctx* ctx, int w, int h, int nfeatures);
/* cv::ORB::detect(img, kps, mask) */
int orbb200_bird_detect(orbb200_ctx* ctx, const uint8_t* img, const uint8_t* mask, int w, int h, size_t stride, size_t mask_stride,
                        int nfeatures, orbb200_kp_t* kps, int cap, int* n_out);
/* cv::cornerSubPix(img, pts, Size(win_w, win_h), Size(-1,-1), TermCriteria(EPS + MAX_ITER, max_iter, eps)); pts [n][2] in/out;
 * window half-sizes 1..7. */
int orbb200_corner_subpix(orbb200_ctx* ctx, const uint8_t* img, int w, int h, size_t stride, float* pts, int n, int win_w, int win_h,
                          int max_iter, double eps);
/* cv::ORB::compute(img, kps, desc): kps [n] in/out (keypoints closer than 31 px to the image border are removed, the
 * rest keep their order), desc [n][32]; *n_out = keypoints left. */
int orbb200_bird_compute(orbb200_ctx* ctx, const uint8_t* img, int w, int h, size_t stride, orbb200_kp_t* kps, int n, uint8_t* desc,
                         int* n_out);
/* detect + cornerSubPix(5,5; 40 it; 1e-3) + compute without leaving the device: mvKeysBird / mDescriptorsBird. */
int orbb200_bird_extract(orbb200_ctx* ctx, const uint8_t* img, const uint8_t* mask, int w, int h, size_t stride, size_t mask_stride,
                         int nfeatures, orbb200_kp_t* kps, uint8_t* desc, int cap, int* n_out);
/* n images (and masks, or masks == NULL) of one size per call; outputs [n][cap_per_img]. */
int orbb200_bird_extract_batch(orbb200_ctx* ctx, const uint8_t* const* imgs, const uint8_t* const* masks, int n, int w, int h,
                               size_t stride, size_t mask_stride, int nfeatures, orbb200_kp_t* kps, uint8_t* desc, int cap_per_img,
                               int* n_out);

/* ---- DBoW2 vocabulary transform ----------------------------------------------------------------------------
 * Frame::ComputeBoW (src/Frame.cc:562-569) -> ORBVocabulary::transform(features, BowVector, FeatureVector, levelsup)
 * (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1139-1203,1230-1271), TF_IDF weighting + L1 scoring.  The tree is
 * passed flattened: children as CSR (child_ptr[n_nodes+1], child_idx), node descriptors [n_nodes][32], and per node
 * the word id (-1 for inner nodes) and weight (idf; 0 = stopped word).  Node 0 is the root.
 * Outputs: per-feature word and node id, the BowVector as ascending (word, value) pairs and the FeatureVector as CSR
 * over ascending node ids -- the form orbb200_search_by_bow / orbb200_search_for_triangulation take. */
typedef struct orbb200_voc orbb200_voc;
int orbb200_voc_create(orbb200_ctx* ctx, orbb200_voc** voc, int n_nodes, const int32_t* child_ptr, const int32_t* child_idx,
                       const uint8_t* node_desc, const int32_t* word_id, const double* weight, int L);
void orbb200_voc_free(orbb200_voc* voc);
int orbb200_bow_transform(orbb200_ctx* ctx, const orbb200_voc* voc, const uint8_t* desc, int n, int levelsup,
                          int32_t* out_word /*[n] or NULL*/, int32_t* out_node /*[n] or NULL*/,
                          int32_t* bow_word /*[n]*/, double* bow_value /*[n]*/, int* n_words,
                          int32_t* fv_node /*[n]*/, int32_t* fv_ptr /*[n+1]*/, int32_t* fv_idx /*[n]*/, int* n_fv);
/* Same on the descriptors of image `img_index` of the last extraction (no upload). */
int orbb200_bow_transform_extracted(orbb200_ctx* ctx, const orbb200_voc* voc, int img_index, int levelsup,
                                    int32_t* out_word, int32_t* out_node, int32_t* bow_word, double* bow_value, int* n_words,
                                    int32_t* fv_node, int32_t* fv_ptr, int32_t* fv_idx, int* n_fv);

/* ---- batched front-end step (bench / sequence processing) -----------------------------------------
 * One pass of the C2 hot path over a batch: extract 2*n_frames images (left,right interleaved: image 2i is
 * the left image of frame i), build the left frame's grid and run SearchByProjection of nq_per_frame
 * queries against it, all on the device.  Query arrays are [n_frames][nq_per_frame] device arrays. */
typedef struct {
    const uint8_t* q_valid; const float* q_u; const float* q_v; const float* q_uR;
    const int32_t* q_level; const float* q_viewcos; const uint8_t* q_desc; const uint8_t* q_obs_pos;
} orbb200_proj_queries;
int orbb200_stereo_step_device(orbb200_ctx* ctx, const uint8_t* d_imgs, size_t img_bytes, int n_frames, int w, int h,
                               size_t stride, const orbb200_proj_queries* d_queries, int nq_per_frame,
                               float th, float nnratio, float min_x, float min_y, float inv_w, float inv_h,
                               int32_t* d_out_best_idx /*[n_frames][nq]*/, int32_t* d_out_best_dist,
                               int32_t* d_nmatches /*[n_frames]*/);

/* Same step with HOST buffers (the end-to-end path): copies the 2*n_frames images (contiguous, rows of
 * `stride` bytes) and the query arrays to the device, runs the step, and copies keypoints, descriptors,
 * counts and match results back.  Asynchronous on the context's stream: outputs are valid after
 * orbb200_sync().  Pinned host memory makes the copies overlap with other contexts' work. */
int orbb200_stereo_step_host(orbb200_ctx* ctx, const uint8_t* h_imgs, int n_frames, int w, int h, size_t stride,
                             const orbb200_proj_queries* h_queries, int nq_per_frame,
                             float th, float nnratio, float min_x, float min_y, float inv_w, float inv_h,
                             orbb200_kp_t* h_kps /*[2n][cap]*/, uint8_t* h_desc /*[2n][cap][32]*/, int cap_per_img,
                             int32_t* h_counts /*[2n]*/, int32_t* h_best_idx /*[n][nq]*/, int32_t* h_best_dist,
                             int32_t* h_nmatches /*[n]*/);

/* Make the batched step also run ComputeStereoMatches between extraction and matching (the left frames then carry
 * mvuRight, which SearchByProjection's stereo-consistency test reads, src/ORBmatcher.cc:91-96). */
int orbb200_step_enable_stereo(orbb200_ctx* ctx, int enable, float mb, float mbf);

/* ---- batched frame step: stereo front camera + birdview (the north-star frame) -------------------------------
 * Everything the data-parallel side of one tracked frame does in the reference, for n_frames frames per call:
 *   Frame::Frame (src/Frame.cc:84-142, 263-375): ORBextractor on the left and right image (:124-127), ComputeStereoMatches
 *     (:662-836), cv::ORB(2000) detect(mask) + cornerSubPix + compute on the birdview image (:328-342), the two lookup grids
 *     (:378-412);
 *   Tracking::SearchLocalPoints (src/Tracking.cc:1610-1660): Frame::isInFrustum for every local map point, then
 *     ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th) (src/ORBmatcher.cc:45-129);
 *   ORBmatcher::SearchByMatchBird(Cur, Last, window) = BirdviewMatch(Last, Cur, vnMatches12, window)
 *     (src/ORBmatcher.cc:1901-1921, 1788-1899) against the previous frame: frame i-1 of the call, or, for frame 0 with
 *     chain != 0, the last frame of the previous call on this context (sequence processing).
 * The local map stays on the device (orbb200_map_upload); a frame sends its pose.  The birdview mask is the constant
 * vehicle mask of the reference (Examples/Monocular/mask_new_front.png is its front-camera analogue): set it once. */
typedef struct {
    int n_frames;
    int w, h; size_t stride;                   /* front camera: 2*n_frames images, image 2i = left, 2i+1 = right of frame i */
    float mb, mbf;                             /* stereo baseline, baseline*fx; mb <= 0: no ComputeStereoMatches (monocular) */
    float min_x, min_y, inv_w, inv_h;          /* front lookup grid (mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv) */
    const orbb200_map* map;                    /* local map or NULL (no SearchLocalPoints) */
    float viewing_cos_limit, th, nnratio;      /* isInFrustum(pMP, 0.5); SearchByProjection(F, points, th) of ORBmatcher(nnratio) */
    int bird_w, bird_h; size_t bird_stride;    /* birdview: n_frames images; bird_w == 0: no birdview */
    int bird_nfeatures;                        /* cv::ORB::create(bird_nfeatures) */
    int bird_window; float bird_nnratio; int bird_check_ori;   /* SearchByMatchBird(Cur, Last, 15) of ORBmatcher(0.99, true) */
    int chain;                                 /* frame 0 continues the sequence of the previous call */
} orbb200_frame_step_params;
typedef struct {                               /* device pointers (_device) or host pointers (_host) */
    const uint8_t* imgs;                       /* [2*n_frames] front images, contiguous, h*stride bytes each */
    const uint8_t* bird_imgs;                  /* [n_frames] birdview images, bird_h*bird_stride bytes each */
    const orbb200_camera_pose* poses;          /* [n_frames] */
} orbb200_frame_step_inputs;
typedef struct {                               /* any pointer may be NULL (result not wanted / stays in the device pools) */
    orbb200_kp_t* kps; uint8_t* desc; int32_t* counts;           /* [2n][cap] / [2n][cap][32] / [2n]   mvKeys, mvKeysRight, mDescriptors[Right] */
    float* u_right; float* depth;                                /* [n][cap]                          mvuRight, mvDepth */
    int32_t* map_best_idx; int32_t* map_best_dist;               /* [n][map_n]  keypoint matched to each map point (-1) and its distance */
    int32_t* map_nmatches;                                       /* [n] */
    orbb200_kp_t* bird_kps; uint8_t* bird_desc; int32_t* bird_counts;   /* [n][bird_cap]                 mvKeysBird, mDescriptorsBird */
    int32_t* bird_matches12; int32_t* bird_nmatches;             /* [n][bird_cap]: vnMatches12 of BirdviewMatch(previous, frame i); [n] */
    int cap, bird_cap;                                           /* row capacities of the arrays above (host variant) */
} orbb200_frame_step_outputs;
/* capacity (row length) of the birdview keypoint pools for this size: bird_cap of the device variant */
int orbb200_bird_set_mask(orbb200_ctx* ctx, int w, int h, int nfeatures, int max_batch, const uint8_t* mask, size_t mask_stride);
/* Device variant: asynchronous on the context's stream; the map_* / bird_matches12 / bird_nmatches arrays are device arrays with
 * rows of map_n / orbb200_bird_max_keypoints() entries; kps/desc/u_right/bird_kps/... are ignored (results stay in the pools:
 * orbb200_results_device, orbb200_stereo_results_device, orbb200_bird_results_device). */
int orbb200_frame_step_device(orbb200_ctx* ctx, const orbb200_frame_step_params* p, const orbb200_frame_step_inputs* d_in,
                              const orbb200_frame_step_outputs* d_out);
/* Host variant (the end-to-end path): copies images and poses to the device, runs the step, copies every non-NULL output back.
 * Asynchronous: outputs are valid after orbb200_sync() (calls of up to 8 frames from pageable memory pass through a pinned staging
 * block and are delivered by orbb200_sync() itself: synchronising the stream alone is not enough for them).  Pinned host memory
 * lets the copies overlap other contexts' work. */
int orbb200_frame_step_host(orbb200_ctx* ctx, const orbb200_frame_step_params* p, const orbb200_frame_step_inputs* h_in,
                            const orbb200_frame_step_outputs* h_out);
int orbb200_bird_results_device(orbb200_ctx* ctx, int w, int h, int nfeatures, const orbb200_kp_t** d_kps, const uint8_t** d_desc,
                                const int32_t** d_counts, int* cap_per_img);
/* Device-side error flags raised since the last call (octree / birdview selection overflow): 0 = none.  Synchronises. */
int orbb200_device_status(orbb200_ctx* ctx, int* status);

/* Per-stage device timing (bench): CUDA events on the context's stream around each stage.
 * stages: 0 import, 1 pyramid, 2 FAST, 3 blur, 4 octree, 5 orientation+descriptors, 6 grid build, 7 windowed match,
 * 8 stereo matching, 9 birdview import+pyramid, 10 birdview FAST+retainBest+Harris+angles, 11 birdview cornerSubPix,
 * 12 birdview blur+descriptors, 13 isInFrustum, 14-15 reserved */
#define ORBB200_NUM_STAGES 16
int orbb200_stage_timing(orbb200_ctx* ctx, int enable);
/* Synchronises, then returns accumulated milliseconds and launch-group counts per stage; reset!=0 clears. */
int orbb200_stage_times(orbb200_ctx* ctx, float* ms /*[ORBB200_NUM_STAGES]*/, int32_t* groups /*[ORBB200_NUM_STAGES]*/, int reset);
/* Measured POPC issue rate of this device in G popc/s (denominator of the matching roofline). */
double orbb200_measure_popc_peak(orbb200_ctx* ctx);

/* Number of kernel launches this context has enqueued since creation (bench bookkeeping). */
long long orbb200_launch_count(const orbb200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
