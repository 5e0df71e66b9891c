// Where one north-star frame (one frame per orbb200_frame_step_host call, results on the host before the next call) spends its
// time, measured from C++ (no interpreter in the loop).  Legs:
//   host_full_ms : orbb200_frame_step_host from pageable memory + orbb200_sync, wall per frame
//   front_only / bird_only : the same call without the birdview image / without map and stereo (which front-end is the long one)
//   dev_events_ms: orbb200_frame_step_device on device-resident inputs, CUDA events around it (the device chain alone)
// Synthetic textured images (a realistic keypoint load, not the bench images), 3000 random map points in front of the camera.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "orbb200.h"

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static unsigned rs = 99991;
static unsigned rnd() { rs = rs * 1664525u + 1013904223u; return rs >> 8; }
static float rndf() { return (float)(rnd() & 0xffff) / 65536.f; }

static void texture(std::vector<uint8_t>& img, int w, int h, int shift)
{
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const int xs = x + shift;
            const int blk = ((xs / 13) * 7 + (y / 11) * 13 + ((xs / 47) ^ (y / 29)) * 3) % 5;
            img[(size_t)y * w + x] = (uint8_t)(40 + blk * 40 + ((xs * 31 + y * 17) & 7));
        }
}

int main(int argc, char** argv)
{
    const int W = 1241, H = 376, BW = 400, BH = 400, NF = 2000, NMAP = 3000;
    const int reps = argc > 1 ? atoi(argv[1]) : 100;
    orbb200_ctx* ctx = nullptr;
    if (orbb200_create(&ctx, 0, NF, 1.2f, 8, 20, 7, W, H, 2) != 0) { fprintf(stderr, "create: %s\n", orbb200_last_error(nullptr)); return 1; }
    const int cap = orbb200_max_keypoints(ctx), bcap = orbb200_bird_max_keypoints(ctx, BW, BH, 2000);
    std::vector<uint8_t> imgs((size_t)2 * W * H), bird((size_t)BW * BH), mask((size_t)BW * BH, 255);
    for (int y = 140; y < 260; y++) for (int x = 160; x < 240; x++) mask[(size_t)y * BW + x] = 0;
    orbb200_bird_set_mask(ctx, BW, BH, 2000, 1, mask.data(), BW);
    // map: points on a plane z = 10 in front of an identity camera
    std::vector<float> pos(3 * NMAP), nrm(3 * NMAP), dmax(NMAP), dmin(NMAP);
    std::vector<uint8_t> mdesc((size_t)NMAP * 32);
    const float fx = 718.856f, fy = 718.856f, cx = 607.19f, cy = 185.2f;
    for (int i = 0; i < NMAP; i++) {
        const float u = rndf() * W, v = rndf() * H, z = 10.f;
        pos[3 * i] = (u - cx) * z / fx; pos[3 * i + 1] = (v - cy) * z / fy; pos[3 * i + 2] = z;
        const float d = std::sqrt(pos[3 * i] * pos[3 * i] + pos[3 * i + 1] * pos[3 * i + 1] + z * z);
        nrm[3 * i] = pos[3 * i] / d; nrm[3 * i + 1] = pos[3 * i + 1] / d; nrm[3 * i + 2] = z / d;
        dmax[i] = d * 2.f; dmin[i] = d * 0.4f;
        for (int k = 0; k < 32; k++) mdesc[(size_t)i * 32 + k] = (uint8_t)rnd();
    }
    orbb200_map* map = nullptr;
    if (orbb200_map_upload(ctx, &map, NMAP, pos.data(), nrm.data(), dmax.data(), dmin.data(), mdesc.data()) != 0) { fprintf(stderr, "map: %s\n", orbb200_last_error(ctx)); return 1; }
    orbb200_camera_pose pose{};
    pose.Rcw[0] = pose.Rcw[4] = pose.Rcw[8] = 1.f;
    pose.fx = fx; pose.fy = fy; pose.cx = cx; pose.cy = cy; pose.mbf = 386.1448f;
    pose.min_x = 0; pose.max_x = (float)W; pose.min_y = 0; pose.max_y = (float)H; pose.log_scale_factor = std::log(1.2f); pose.n_levels = 8;
    orbb200_frame_step_params P{};
    P.n_frames = 1; P.w = W; P.h = H; P.stride = W; P.mb = 0.537f; P.mbf = 386.1448f;
    P.min_x = 0; P.min_y = 0; P.inv_w = 64.f / W; P.inv_h = 48.f / H; P.map = map; P.viewing_cos_limit = 0.5f; P.th = 1.f; P.nnratio = 0.8f;
    P.bird_w = BW; P.bird_h = BH; P.bird_stride = BW; P.bird_nfeatures = 2000; P.bird_window = 15; P.bird_nnratio = 0.99f; P.bird_check_ori = 1;
    std::vector<orbb200_kp_t> kps((size_t)2 * cap), bkps(bcap);
    std::vector<uint8_t> desc((size_t)2 * cap * 32), bdesc((size_t)bcap * 32);
    std::vector<float> ur(cap), dep(cap);
    std::vector<int32_t> counts(2), bi(NMAP), bd(NMAP), nm(1), bcnt(1), m12(bcap), bnm(1);
    orbb200_frame_step_outputs O{};
    O.kps = kps.data(); O.desc = desc.data(); O.counts = counts.data(); O.u_right = ur.data(); O.depth = dep.data();
    O.map_best_idx = bi.data(); O.map_best_dist = bd.data(); O.map_nmatches = nm.data();
    O.bird_kps = bkps.data(); O.bird_desc = bdesc.data(); O.bird_counts = bcnt.data(); O.bird_matches12 = m12.data(); O.bird_nmatches = bnm.data();
    O.cap = cap; O.bird_cap = bcap;
    orbb200_frame_step_inputs I{};
    I.imgs = imgs.data(); I.bird_imgs = bird.data(); I.poses = &pose;
    auto frame = [&](int i, const orbb200_frame_step_params& PP, const orbb200_frame_step_outputs& OO) {
        std::vector<uint8_t> l((size_t)W * H), r((size_t)W * H);
        (void)l; (void)r;
        P.chain = i > 0;
        orbb200_frame_step_params Q = PP; Q.chain = i > 0;
        if (orbb200_frame_step_host(ctx, &Q, &I, &OO) != 0 || orbb200_sync(ctx) != 0) { fprintf(stderr, "frame_step: %s\n", orbb200_last_error(ctx)); exit(1); }
    };
    // a short sequence of distinct frames, cycled (the content changes from call to call like a real sequence)
    const int NSEQ = 6;
    std::vector<std::vector<uint8_t>> seqI(NSEQ, std::vector<uint8_t>((size_t)2 * W * H)), seqB(NSEQ, std::vector<uint8_t>((size_t)BW * BH));
    for (int s = 0; s < NSEQ; s++) {
        std::vector<uint8_t> a((size_t)W * H), b((size_t)W * H);
        texture(a, W, H, 3 * s); texture(b, W, H, 3 * s + 7);
        memcpy(seqI[s].data(), a.data(), a.size()); memcpy(seqI[s].data() + a.size(), b.data(), b.size());
        texture(seqB[s], BW, BH, 2 * s);
    }
    auto run = [&](const orbb200_frame_step_params& PP, const orbb200_frame_step_outputs& OO) {
        for (int i = 0; i < 6; i++) { I.imgs = seqI[i % NSEQ].data(); I.bird_imgs = seqB[i % NSEQ].data(); frame(i, PP, OO); }
        const double t0 = now_ms();
        for (int i = 0; i < reps; i++) { I.imgs = seqI[i % NSEQ].data(); I.bird_imgs = seqB[i % NSEQ].data(); frame(6 + i, PP, OO); }
        return (now_ms() - t0) / reps;
    };
    const double full = run(P, O);
    const int kpL = counts[0], kpB = bcnt[0], nmatch = nm[0], nb = bnm[0];
    orbb200_frame_step_params Pf = P; Pf.bird_w = 0; Pf.bird_h = 0;
    orbb200_frame_step_outputs Of = O; Of.bird_kps = nullptr; Of.bird_desc = nullptr; Of.bird_counts = nullptr; Of.bird_matches12 = nullptr; Of.bird_nmatches = nullptr;
    const double front = run(Pf, Of);
    orbb200_frame_step_params Pb = P; Pb.map = nullptr; Pb.mb = 0.f;
    orbb200_frame_step_outputs Ob = O; Ob.map_best_idx = nullptr; Ob.map_best_dist = nullptr; Ob.map_nmatches = nullptr; Ob.u_right = nullptr; Ob.depth = nullptr;
    const double birdish = run(Pb, Ob);
    // device chain alone
    uint8_t *dI = nullptr, *dB = nullptr; orbb200_camera_pose* dP = nullptr; int32_t *dBi, *dBd, *dNm, *dM12, *dBnm;
    cudaMalloc(&dI, imgs.size()); cudaMalloc(&dB, bird.size()); cudaMalloc(&dP, sizeof(pose));
    cudaMalloc(&dBi, 4 * NMAP); cudaMalloc(&dBd, 4 * NMAP); cudaMalloc(&dNm, 4); cudaMalloc(&dM12, 4 * (size_t)bcap); cudaMalloc(&dBnm, 4);
    cudaMemcpy(dI, seqI[0].data(), imgs.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, seqB[0].data(), bird.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dP, &pose, sizeof(pose), cudaMemcpyHostToDevice);
    orbb200_frame_step_inputs dIn{dI, dB, dP};
    orbb200_frame_step_outputs dOut{};
    dOut.map_best_idx = dBi; dOut.map_best_dist = dBd; dOut.map_nmatches = dNm; dOut.bird_matches12 = dM12; dOut.bird_nmatches = dBnm;
    cudaStream_t st = (cudaStream_t)orbb200_stream(ctx);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    orbb200_frame_step_params Q = P; Q.chain = 1;
    double dev = 0;
    for (int i = 0; i < reps + 3; i++) {
        cudaEventRecord(e0, st);
        if (orbb200_frame_step_device(ctx, &Q, &dIn, &dOut) != 0) { fprintf(stderr, "frame_step_device: %s\n", orbb200_last_error(ctx)); return 1; }
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (i >= 3) dev += ms;
    }
    printf("{\"host_full_ms\": %.4f, \"front_only_ms\": %.4f, \"bird_only_ms\": %.4f, \"dev_events_ms\": %.4f, \"kps_left\": %d, \"kps_bird\": %d, "
           "\"map_matches\": %d, \"bird_matches\": %d}\n", full, front, birdish, dev / reps, kpL, kpB, nmatch, nb);
    orbb200_map_free(map);
    orbb200_destroy(ctx);
    return 0;
}
