// ORACLE -- test infrastructure only.  NOT part of the product and never linked into liborbb200.so.
//
// The matcher entry points of include/orbb200.h implemented on the CPU oracle (oracle/match_oracle.cpp), so that the C++
// adapters of orb-slam-birdview_b200/cpp/ORBmatcher_b200.cc can be linked into a CPU-only test binary
// (oracle/_ref/matcher_suite_oracle) and compared, in a container without a GPU, with the reference's own unmodified
// src/ORBmatcher.cc (oracle/_ref/matcher_suite_ref).  That comparison pins two things at once: the oracle's matcher loops on
// the real reference code, and the host-side geometry of the adapters.  The product binary (cpp/matcher_suite) links the same
// adapters against liborbb200.so and needs a GPU; tests/test_matcher_suite.py runs all three.
#include <cstring>
#include <string>
#include <vector>

#include "../include/orbb200.h"
#include "orb_oracle.h"

struct orbb200_ctx { std::vector<float> sf, isf, s2, is2; };
struct orbb200_frame { oracle_frame* f; int n; };

extern "C" {

int orbb200_create(orbb200_ctx** out, int, int, float scaleFactor, int nlevels, int, int, int, int, int)
{
    orbb200_ctx* c = new orbb200_ctx;
    c->sf.resize(nlevels); c->isf.resize(nlevels); c->s2.resize(nlevels); c->is2.resize(nlevels);
    c->sf[0] = 1.0f; c->s2[0] = 1.0f;                                   // src/ORBextractor.cc:415-426
    for (int i = 1; i < nlevels; i++) { c->sf[i] = c->sf[i - 1] * scaleFactor; c->s2[i] = c->sf[i] * c->sf[i]; }
    for (int i = 0; i < nlevels; i++) { c->isf[i] = 1.0f / c->sf[i]; c->is2[i] = 1.0f / c->s2[i]; }
    *out = c;
    return ORBB200_OK;
}
void orbb200_destroy(orbb200_ctx* c) { delete c; }
const char* orbb200_last_error(const orbb200_ctx*) { return "oracle-backed test ABI"; }
int orbb200_get_levels(const orbb200_ctx* c) { return (int)c->sf.size(); }
int orbb200_get_scale_table(const orbb200_ctx* c, int which, float* out)
{
    const std::vector<float>& t = which == 0 ? c->sf : which == 1 ? c->isf : which == 2 ? c->s2 : c->is2;
    memcpy(out, t.data(), t.size() * sizeof(float));
    return ORBB200_OK;
}
int orbb200_frame_upload(orbb200_ctx*, orbb200_frame** f, const orbb200_kp_t* kps, const uint8_t* desc, const float* u_right, int n,
                         float min_x, float min_y, float inv_w, float inv_h)
{
    *f = new orbb200_frame{oracle_frame_create((const oracle_kp_t*)kps, desc, n, min_x, min_y, inv_w, inv_h, u_right), n};
    return ORBB200_OK;
}
void orbb200_frame_free(orbb200_frame* f) { if (f) { oracle_frame_destroy(f->f); delete f; } }

int orbb200_search_by_projection(orbb200_ctx* c, const orbb200_frame* F, int nq, const uint8_t* q_valid, const float* q_u, const float* q_v,
                                 const float* q_uR, const int32_t* q_level, const float* q_viewcos, const uint8_t* q_desc, const uint8_t* q_obs_pos,
                                 const uint8_t* kp_blocked, float th, float nnratio, int32_t* bi, int32_t* bd, int32_t* qk, int* nmatches)
{
    *nmatches = oracle_search_by_projection(F->f, c->sf.data(), nq, q_valid, q_u, q_v, q_uR, q_level, q_viewcos, q_desc, q_obs_pos, kp_blocked, th, nnratio, bi, bd, qk);
    return ORBB200_OK;
}
int orbb200_search_by_projection_frame(orbb200_ctx* c, const orbb200_frame* Cur, int nq, const uint8_t* q_valid, const float* q_u, const float* q_v,
                                       const float* q_invz, const int32_t* q_octave, const float* q_angle, const uint8_t* q_desc, const uint8_t* q_obs_pos,
                                       const uint8_t* kp_blocked, float th, float mbf, int mode, int check_ori, int32_t* qk, int* nmatches)
{
    *nmatches = oracle_search_by_projection_frame(Cur->f, c->sf.data(), nq, q_valid, q_u, q_v, q_invz, q_octave, q_angle, q_desc, q_obs_pos, kp_blocked, th, mbf,
                                                  mode, check_ori, qk);
    return ORBB200_OK;
}
int orbb200_birdview_match(orbb200_ctx*, const orbb200_kp_t* kps1, const uint8_t* desc1, int n1, const orbb200_frame* F2, float* prev_xy, int window_size,
                           float nnratio, int check_ori, int32_t* matches12, int* nmatches)
{
    *nmatches = oracle_birdview_match((const oracle_kp_t*)kps1, desc1, n1, F2->f, prev_xy, window_size, nnratio, check_ori, matches12);
    return ORBB200_OK;
}
int orbb200_search_by_match_bird_kf(orbb200_ctx*, const orbb200_kp_t* kf_kps, const uint8_t* has_mp, const uint8_t* mp_desc, int nk, const orbb200_frame* F,
                                    float r, float nnratio, int check_ori, int32_t* out_mp_of_kp, int* nmatches)
{
    *nmatches = oracle_search_by_match_bird_kf((const oracle_kp_t*)kf_kps, has_mp, mp_desc, nk, F->f, r, nnratio, check_ori, out_mp_of_kp);
    return ORBB200_OK;
}
int orbb200_search_by_projection_bird(orbb200_ctx*, const orbb200_frame* F, int nq, const uint8_t* q_valid, const float* q_x, const float* q_y,
                                      const uint8_t* q_desc, const uint8_t* q_obs_pos, const uint8_t* kp_blocked, float r, float nnratio, int32_t* qk, int* nmatches)
{
    *nmatches = oracle_search_by_projection_bird(F->f, nq, q_valid, q_x, q_y, q_desc, q_obs_pos, kp_blocked, r, nnratio, qk);
    return ORBB200_OK;
}
int orbb200_search_for_triangulation(orbb200_ctx*, const orbb200_kp_t* kps1, const uint8_t* desc1, const float* uR1, const uint8_t* has_mp1, int n1,
                                     const orbb200_kp_t* kps2, const uint8_t* desc2, const float* uR2, const uint8_t* has_mp2, int n2,
                                     const int32_t* fv1_node, const int32_t* fv1_ptr, const int32_t* fv1_idx, int nn1,
                                     const int32_t* fv2_node, const int32_t* fv2_ptr, const int32_t* fv2_idx, int nn2,
                                     const float* F12, float ex, float ey, const float* scale_factors2, const float* level_sigma2_2,
                                     int only_stereo, int check_ori, int32_t* pairs, int* npairs)
{
    *npairs = oracle_search_for_triangulation((const oracle_kp_t*)kps1, desc1, uR1, has_mp1, n1, (const oracle_kp_t*)kps2, desc2, uR2, has_mp2, n2,
                                              fv1_node, fv1_ptr, fv1_idx, nn1, fv2_node, fv2_ptr, fv2_idx, nn2, F12, ex, ey, scale_factors2, level_sigma2_2,
                                              only_stereo, check_ori, pairs);
    return ORBB200_OK;
}
int orbb200_search_for_initialization(orbb200_ctx*, const orbb200_kp_t* kps1, const uint8_t* desc1, int n1, const orbb200_frame* F2, float* prev_xy,
                                      int window_size, float nnratio, int check_ori, int32_t* matches12, int* nmatches)
{
    *nmatches = oracle_search_for_initialization((const oracle_kp_t*)kps1, desc1, n1, F2->f, prev_xy, window_size, nnratio, check_ori, matches12);
    return ORBB200_OK;
}
int orbb200_search_window_best(orbb200_ctx*, const orbb200_frame* F, int nq, const uint8_t* q_valid, const float* q_x, const float* q_y, const float* q_r,
                               const int32_t* q_min_level, const int32_t* q_max_level, const uint8_t* q_desc, const float* q_aux, const float* q_angle,
                               const uint8_t* q_obs_pos, const uint8_t* kp_blocked, const float* inv_level_sigma2, int acc_th, int flags,
                               int32_t* bi, int32_t* bd, int32_t* qk, int* nmatches)
{
    *nmatches = oracle_search_window_best(F->f, nq, q_valid, q_x, q_y, q_r, q_min_level, q_max_level, q_desc, q_aux, q_angle, q_obs_pos, kp_blocked,
                                          inv_level_sigma2, acc_th, flags, bi, bd, qk);
    return ORBB200_OK;
}
int orbb200_search_by_bow(orbb200_ctx*, const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1, const orbb200_frame* F2, const uint8_t* valid2,
                          const int32_t* fv1_node, const int32_t* fv1_ptr, const int32_t* fv1_idx, int nn1,
                          const int32_t* fv2_node, const int32_t* fv2_ptr, const int32_t* fv2_idx, int nn2,
                          float nnratio, int check_ori, int kf_kf, int32_t* out, int* nmatches)
{
    *nmatches = oracle_search_by_bow(desc1, angle1, valid1, n1, F2->f, valid2, fv1_node, fv1_ptr, fv1_idx, nn1, fv2_node, fv2_ptr, fv2_idx, nn2, nnratio, check_ori,
                                     kf_kf, out);
    return ORBB200_OK;
}

}  // extern "C"
