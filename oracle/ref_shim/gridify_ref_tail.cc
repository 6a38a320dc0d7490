// gridify_ref_tail.cc -- TEST INFRASTRUCTURE: C entry points around the reference kernel bodies that the
// Makefile pipes in front of this file (namespaces ref_g = gridify.cu:102-291, ref_k = gridifyknn.cu:115-333,
// ref_u = gridify_up.cu:102-225).  Each function restates the HOST side of the reference launcher it names
// (scratch allocation + memsets + the two launches), with the CUDA threads run sequentially (cuda_seq.h), and
// the output initialisation of the operator's Forward.  `seconds` is the value the reference reads from
// gettimeofday().tv_usec (gridify.cu:377-379): it only matters when a time-seeded reservoir fires (more than
// max_p_grid points in one voxel, more than max_o_grid occupied voxels).

namespace {
struct Scratch {  // gridify.cu:356-367
    float *locxyzw;
    int *pntidx, *counter, *voxelidx, *vox2coor, *voxcounter, *centcount;
    Scratch(int B, int G, int O, int P) {
        locxyzw = (float *)calloc((size_t)B * G * 4, sizeof(float));
        pntidx = (int *)calloc((size_t)B * G * P, sizeof(int));  // cudaMalloc, not cleared: never read unwritten
        counter = (int *)calloc((size_t)B * G, sizeof(int));
        voxelidx = (int *)malloc((size_t)B * G * sizeof(int));
        memset(voxelidx, 0xff, (size_t)B * G * sizeof(int));
        vox2coor = (int *)calloc((size_t)B * O, sizeof(int));
        voxcounter = (int *)calloc((size_t)B * O, sizeof(int));
        centcount = (int *)calloc((size_t)B, sizeof(int));
    }
    ~Scratch() {
        free(locxyzw); free(pntidx); free(counter); free(voxelidx); free(vox2coor); free(voxcounter); free(centcount);
    }
};

void init_outputs(int B, int O, int P, int *nebidx, float *nebmsk, float *cent, float *centmsk, int *centnum) {
    // GridifyOp::Forward, gridify-inl.h:117-121
    for (long long i = 0; i < (long long)B * O * P; i++) { nebidx[i] = 0; nebmsk[i] = 0.f; }
    for (long long i = 0; i < (long long)B * O * 4; i++) cent[i] = 1.f;
    for (long long i = 0; i < (long long)B * O; i++) centmsk[i] = 0.f;
    for (int i = 0; i < B; i++) centnum[i] = 0;
}
}  // namespace

extern "C" void ref_gridify(const float *data, const int *npts, int B, int N, int O, int P, int ks, int stride,
                            int loc, const float *shift, const float *voxel, const int *grid,
                            unsigned long seconds, int *nebidx, float *nebmsk, float *cent, float *centmsk,
                            int *centnum) {
    const float gridf[3] = {(float)grid[0], (float)grid[1], (float)grid[2]};  // gridify.cu:343-347
    const int G = grid[0] * grid[1] * grid[2], size = ks * ks * ks;
    init_outputs(B, O, P, nebidx, nebmsk, cent, centmsk, centnum);
    Scratch s(B, G, O, P);
    seq_launch((long long)B * N, [&] {  // gridify.cu:369-385
        ref_g::gridify_kernel_build_index<float>(nebidx, nebmsk, cent, centmsk, centnum, s.centcount, data, npts, B, N,
                                                 O, P, ks, stride, loc, shift, voxel, gridf, G, size, s.voxelidx,
                                                 s.vox2coor, s.pntidx, s.locxyzw, s.counter, seconds);
    });
    seq_launch((long long)B * O, [&] {  // gridify.cu:389-397
        ref_g::gridify_kernel_query_neighs<float>(nebidx, nebmsk, cent, centmsk, centnum, data, npts, B, N, O, P, ks,
                                                  stride, loc, shift, voxel, gridf, G, size, s.voxelidx, s.vox2coor,
                                                  s.pntidx, s.locxyzw, s.counter, s.voxcounter, seconds);
    });
}

extern "C" void ref_gridify_knn(const float *data, const int *npts, int B, int N, int O, int P, int ks,
                                int stride, int loc, const float *shift, const float *voxel, const int *grid,
                                unsigned long seconds, int *nebidx, float *nebmsk, float *cent, float *centmsk,
                                int *centnum) {
    const float gridf[3] = {(float)grid[0], (float)grid[1], (float)grid[2]};
    const int G = grid[0] * grid[1] * grid[2], size = ks * ks * ks;
    init_outputs(B, O, P, nebidx, nebmsk, cent, centmsk, centnum);
    Scratch s(B, G, O, P);
    seq_launch((long long)B * N, [&] {  // gridifyknn.cu:411-427
        ref_k::gridifyKNN_kernel_build_index<float>(nebidx, nebmsk, cent, centmsk, centnum, s.centcount, data, npts, B,
                                                    N, O, P, ks, stride, loc, shift, voxel, gridf, G, size, s.voxelidx,
                                                    s.vox2coor, s.pntidx, s.locxyzw, s.counter, seconds);
    });
    seq_launch((long long)B * O, [&] {  // gridifyknn.cu:431-439
        ref_k::gridifyKNN_kernel_query_neighs<float>(nebidx, nebmsk, cent, centmsk, centnum, data, npts, B, N, O, P, ks,
                                                     stride, loc, shift, voxel, gridf, G, size, s.voxelidx, s.vox2coor,
                                                     s.pntidx, s.locxyzw, s.counter, s.voxcounter, seconds);
    });
}

extern "C" void ref_gridify_up(const float *downdata, const float *updata, const int *downnum, const int *upnum,
                               int B, int N, int O, int P, int ks, const float *shift, const float *voxel,
                               const int *grid, unsigned long seconds, int *nebidx, float *nebmsk) {
    const float gridf[3] = {(float)grid[0], (float)grid[1], (float)grid[2]};
    const int G = grid[0] * grid[1] * grid[2], size = ks * ks * ks;
    for (long long i = 0; i < (long long)B * O * P; i++) { nebidx[i] = 0; nebmsk[i] = 0.f; }  // gridify_up-inl.h:111-112
    int *bucket = (int *)calloc((size_t)B * G * P, sizeof(int));   // gridify_up.cu:290-291 (both cleared)
    int *counter = (int *)calloc((size_t)B * G, sizeof(int));
    seq_launch((long long)B * N * size, [&] {  // gridify_up.cu:294-306
        ref_u::gridify_kernel_build_index<float>(nebidx, nebmsk, downdata, updata, downnum, upnum, B, N, O, P, ks,
                                                 shift, voxel, gridf, G, size, bucket, counter, seconds);
    });
    seq_launch((long long)B * O, [&] {  // gridify_up.cu:308-318
        ref_u::gridify_kernel_query_neighs<float>(nebidx, nebmsk, downdata, updata, downnum, upnum, B, N, O, P, ks,
                                                  shift, voxel, gridf, G, size, bucket, counter);
    });
    free(bucket);
    free(counter);
}
