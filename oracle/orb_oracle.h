/* ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the ORB front-end of donglinb/ORB-SLAM-BIRDVIEW (ORBextractor, ORBmatcher
 * Hamming scans, Frame grid lookups).  Dependency-free C++ (no OpenCV): every OpenCV primitive the
 * reference calls is restated here and pinned bit-exactly against cv2 4.13 by tests/test_oracle_cv2.py
 * and the committed fixtures under tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
 * this library; the product (orb-slam-birdview_b200/csrc) never links or calls it.
 *
 * Parity status: the reference ships no tests, golden vectors or fixtures.  The pins are
 * (0) EXTRACTOR: the reference's own src/ORBextractor.cc compiled unmodified (oracle/_ref, built by
 *     oracle/ref_standin/Makefile against a stand-in for the OpenCV container types; its arithmetic primitives are the
 *     cv2-pinned ones below) -- with a monotone allocator its output equals this oracle byte for byte on every tested
 *     input (tests/test_oracle_ref.py);
 * (1) each OpenCV primitive vs cv2 4.13, (2) the full extractor vs an independent Python composition of cv2
 * primitives (tests/ref_py), (3) matchers vs literal Python transcriptions (the matcher sources need Eigen/DBoW2/g2o
 * headers plus the whole Frame/KeyFrame/MapPoint graph: see DESIGN.md section 2 for what is compiled from them).
 *
 * Documented divergences from "the reference binary":
 *   - DistributeOctTree sorts pair<int,ExtractorNode*> (src/ORBextractor.cc:684): size ties are broken
 *     by heap address.  Oracle rule: among equal sizes the most recently created node is expanded first
 *     (what monotonically increasing allocation addresses give: proven by the "bump" build of oracle/_ref).
 *     Under glibc malloc the real binary reuses freed list nodes (tcache), its tie order follows the heap's history
 *     and its output is not even a function of the image: two runs on the same image share 98-99 % of their keypoints,
 *     the same figure as reference-vs-oracle (profiles/r2_ref_tie_report.json).
 *   - Float expressions are evaluated without FMA contraction (-ffp-contract=off); a contracting build of the
 *     reference differs in ~1e-5 of the descriptors (same report).
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* == cv::KeyPoint (28 bytes) */
typedef struct { float x, y, size, angle, response; int32_t octave, class_id; } oracle_kp_t;

/* ---- OpenCV primitives (pinned to cv2 4.13) ---------------------------------------------------- */
/* cv::resize(src,dst,Size(dw,dh),0,0,INTER_LINEAR) on CV_8UC1 */
void oracle_resize_u8(const uint8_t* src, int sw, int sh, size_t sstep, uint8_t* dst, int dw, int dh, size_t dstep);
/* cv::GaussianBlur(src,dst,Size(7,7),2,2,BORDER_REFLECT_101) on CV_8UC1 */
void oracle_gauss7_u8(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep);
/* cv::FAST(img,kps,threshold,true) TYPE_9_16; returns count, writes up to cap (x,y,response) triples */
int oracle_fast9(const uint8_t* img, int w, int h, size_t step, int threshold, int nms, int32_t* xyr, int cap);
/* threshold-independent FAST score map (M-1, 0 where undefined); for debugging GPU stages */
void oracle_fast_score_map(const uint8_t* img, int w, int h, size_t step, int32_t* score);
float oracle_fast_atan2(float y, float x);
int oracle_cv_round(float v);

/* ---- ORBextractor ------------------------------------------------------------------------------ */
typedef struct oracle_extractor oracle_extractor;
oracle_extractor* oracle_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
void oracle_extractor_destroy(oracle_extractor*);
/* ORBextractor::operator(); returns N (number of keypoints), or -1 if cap too small */
int oracle_extract(oracle_extractor*, const uint8_t* img, int w, int h, size_t step,
                   oracle_kp_t* kps, uint8_t* desc, int cap);
/* accumulated wall milliseconds per stage (pyramid, FAST cells, octree, orientation, blur, descriptors); reset != 0 clears */
void oracle_extractor_stage_ms(oracle_extractor*, double* out6, int reset);
/* tables */
int oracle_extractor_features_per_level(oracle_extractor*, int32_t* out);
int oracle_extractor_scale_factors(oracle_extractor*, float* out);
int oracle_extractor_umax(oracle_extractor*, int32_t* out);
/* intermediates of the last oracle_extract call */
int oracle_extractor_level_size(oracle_extractor*, int level, int* w, int* h);
int oracle_extractor_level_image(oracle_extractor*, int level, int blurred, uint8_t* dst, size_t dstep);
/* FAST candidates of a level fed to DistributeOctTree, region coordinates, in reference order */
int oracle_extractor_level_candidates(oracle_extractor*, int level, int32_t* xyr, int cap);
/* keypoints per level after octree+orientation, level coordinates (before pt*=scale) */
int oracle_extractor_level_keypoints(oracle_extractor*, int level, oracle_kp_t* kps, int cap);
/* Frame::ComputeStereoMatches (src/Frame.cc:662-836): EL/ER = extractors after operator() on the left/right image
 * (their mvImagePyramid is read); returns the number of stereo matches kept, writes mvuRight / mvDepth */
int oracle_compute_stereo_matches(oracle_extractor* EL, oracle_extractor* ER,
                                  const oracle_kp_t* kpsL, const uint8_t* descL, int nL,
                                  const oracle_kp_t* kpsR, const uint8_t* descR, int nR,
                                  float mb, float mbf, float* uRight, float* depth);
/* standalone DistributeOctTree on (x,y,response) candidates in region coords */
int oracle_distribute_octree(const int32_t* xyr, int n, int minX, int maxX, int minY, int maxY, int N,
                             int32_t* out_xyr, int cap);

/* ---- ORBmatcher / Frame grid ------------------------------------------------------------------- */
int oracle_descriptor_distance(const uint8_t* a, const uint8_t* b);
/* brute force best / second best (strict '<': first minimum in scan order wins) */
void oracle_distinctive_descriptors(const uint8_t* desc, const int32_t* group_ptr, int n_groups,
                                    int32_t* best_idx, int32_t* best_median);
void oracle_hamming_knn2(const uint8_t* q, int nq, const uint8_t* m, int nm,
                         int32_t* best_idx, int32_t* best_d, int32_t* second_d);

typedef struct oracle_frame oracle_frame;
/* Frame lookup grid (64x48): Frame::AssignFeaturesToGrid / PosInGrid[Birdview] */
oracle_frame* oracle_frame_create(const oracle_kp_t* kps, const uint8_t* desc, int n,
                                  float minX, float minY, float invW, float invH,
                                  const float* uRight /* may be NULL -> -1 */);
void oracle_frame_destroy(oracle_frame*);
/* Frame::GetFeaturesInArea[Birdview]; returns count */
int oracle_frame_features_in_area(const oracle_frame*, float x, float y, float r, int minLevel, int maxLevel,
                                  int32_t* out, int cap);

/* ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th)  (src/ORBmatcher.cc:45-129)
 * queries: per map point  valid (mbTrackInView && !isBad), u,v,uR (mTrackProj*), level, viewCos, desc,
 * obs_pos (Observations()>0 of the point, governs blocking after assignment).
 * kp_blocked: per keypoint 1 if it already holds a map point with Observations()>0.
 * out_query_of_kp[n]: index of the query finally assigned to each keypoint (-1 none).  returns nmatches */
int oracle_search_by_projection(const oracle_frame* F, const float* scaleFactors, int nq,
                                const uint8_t* q_valid, const float* q_u, const float* q_v, const float* q_uR,
                                const int32_t* q_level, const float* q_viewcos, const uint8_t* q_desc,
                                const uint8_t* q_obs_pos, const uint8_t* kp_blocked,
                                float th, float nnratio,
                                int32_t* out_best_idx /*[nq]*/, int32_t* out_best_dist /*[nq]*/,
                                int32_t* out_query_of_kp /*[n]*/);

/* ORBmatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, bMono) (:1328-1470), after the host
 * has projected Last's map points: per Last keypoint i  valid, u, v, invzc, octave, angle, desc, obs_pos.
 * mode: 0 = levels [oct-1,oct+1], 1 = forward (>=oct), 2 = backward ([0,oct]). */
int oracle_search_by_projection_frame(const oracle_frame* Cur, const float* scaleFactors, int nq,
                                      const uint8_t* q_valid, const float* q_u, const float* q_v,
                                      const float* q_invz, const int32_t* q_octave, const float* q_angle,
                                      const uint8_t* q_desc, const uint8_t* q_obs_pos,
                                      const uint8_t* kp_blocked, float th, float mbf, int mode, int checkOri,
                                      int32_t* out_query_of_kp /*[n]*/);

/* ORBmatcher::BirdviewMatch(F1,F2,vnMatches12,vPrevMatched,win) (:1667-1786) when prev!=NULL (octave-0
 * queries only, window centred on prev[i1]; prev is updated), and
 * ORBmatcher::BirdviewMatch(const F1,const F2,vnMatches12,win) (:1788-1899) when prev==NULL. */
int oracle_birdview_match(const oracle_kp_t* kps1, const uint8_t* desc1, int n1, const oracle_frame* F2,
                          float* prev_xy /* [n1][2] or NULL */, int windowSize, float nnratio, int checkOri,
                          int32_t* matches12 /*[n1]*/);

/* ORBmatcher::SearchByMatchBird(KeyFrame*,Frame&,out,r) (:2000-2114): per KF bird keypoint k with a landmark
 * (has_mp[k]) and that landmark's descriptor mp_desc[k]; out_mp_of_kp[n] = k assigned to F's keypoint or -1 */
int oracle_search_by_match_bird_kf(const oracle_kp_t* kf_kps, const uint8_t* has_mp, const uint8_t* mp_desc, int nk,
                                   const oracle_frame* F, float r, float nnratio, int checkOri,
                                   int32_t* out_mp_of_kp /*[n]*/);

/* ORBmatcher::SearchByProjectionBird(F,vpMapPointsBird,r) (:1923-1998) after projection on the host:
 * per landmark valid (not seen this frame, |z|<=0.2, inside image), pt.x, pt.y, desc, obs_pos. */
int oracle_search_by_projection_bird(const oracle_frame* F, int nq, const uint8_t* q_valid,
                                     const float* q_x, const float* q_y, const uint8_t* q_desc,
                                     const uint8_t* q_obs_pos, const uint8_t* kp_blocked, float r, float nnratio,
                                     int32_t* out_query_of_kp /*[n]*/);

/* ORBmatcher::SearchForInitialization (:405-520) */
int oracle_search_for_initialization(const oracle_kp_t* kps1, const uint8_t* desc1, int n1, const oracle_frame* F2,
                                     float* prev_xy, int windowSize, float nnratio, int checkOri, int32_t* vnMatches12);
/* common loop of SearchByProjection(Frame&,KeyFrame*,set,..) (:1472-1599), SearchByProjection(KeyFrame*,Scw,..) (:290-403),
 * Fuse (:825-1100), SearchBySim3 directions (:1102-1326) after the host geometry; flags: 1 BLOCK, 2 URCHECK, 4 CHI2, 8 ORI */
int oracle_search_window_best(const oracle_frame* F, int nq, const uint8_t* q_valid, const float* q_x, const float* q_y,
                              const float* q_r, const int32_t* q_minL, const int32_t* q_maxL, const uint8_t* q_desc,
                              const float* q_aux, const float* q_angle, const uint8_t* q_obs_pos, const uint8_t* kp_blocked,
                              const float* invLevelSigma2, int accTh, int flags,
                              int32_t* out_best_idx, int32_t* out_best_dist, int32_t* out_query_of_kp);
/* ORBmatcher::SearchByBoW(KeyFrame*,Frame&,..) (:159-288; kf_kf == 0) / (KeyFrame*,KeyFrame*,..) (:522-655; kf_kf != 0) */
int oracle_search_by_bow(const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1,
                         const oracle_frame* F2, const uint8_t* valid2,
                         const int32_t* fv1_node, const int32_t* fv1_ptr, const int32_t* fv1_idx, int nn1,
                         const int32_t* fv2_node, const int32_t* fv2_ptr, const int32_t* fv2_idx, int nn2,
                         float nnratio, int checkOri, int kf_kf, int32_t* out);

/* ORBmatcher::SearchForTriangulation(KF1,KF2,F12,pairs,bOnlyStereo) (:657-823).
 * Feature vectors are given as CSR over shared-vocabulary node ids sorted ascending:
 * fv?_node[nn?], fv?_ptr[nn?+1], fv?_idx[...].  has_mp?: keypoint already has a MapPoint.
 * Returns nmatches; pairs written as (idx1, idx2). */
int oracle_search_for_triangulation(const oracle_kp_t* kps1, const uint8_t* desc1, const float* uR1, const uint8_t* has_mp1, int n1,
                                    const oracle_kp_t* kps2, const uint8_t* desc2, const float* uR2, const uint8_t* has_mp2, int n2,
                                    const int32_t* fv1_node, const int32_t* fv1_ptr, const int32_t* fv1_idx, int nn1,
                                    const int32_t* fv2_node, const int32_t* fv2_ptr, const int32_t* fv2_idx, int nn2,
                                    const float* F12 /*3x3 row-major*/, float ex, float ey,
                                    const float* scaleFactors2, const float* levelSigma2_2,
                                    int onlyStereo, int checkOri,
                                    int32_t* pairs /*[n1][2]*/);

/* ---- DBoW2 transform (TemplatedVocabulary.h:1139-1271), TF_IDF + L1 ------------------------------------- */
typedef struct oracle_voc oracle_voc;
oracle_voc* oracle_voc_create(int nNodes, const int32_t* child_ptr, const int32_t* child_idx, const uint8_t* node_desc,
                              const int32_t* word_id, const double* weight, int L);
void oracle_voc_destroy(oracle_voc*);
int oracle_voc_transform(const oracle_voc* V, const uint8_t* features, int n, int levelsup,
                         int32_t* out_word, int32_t* out_node,
                         int32_t* bow_word, double* bow_value, int32_t* n_words,
                         int32_t* fv_node, int32_t* fv_ptr, int32_t* fv_idx, int32_t* n_fv);

#ifdef __cplusplus
}
#endif
#endif
