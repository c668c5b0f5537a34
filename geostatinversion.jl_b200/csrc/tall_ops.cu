// Layout plumbing for device buffers: host column-major <-> device TALL (row-major,
// pitch ld) conversion through a bounded staging area, COLMAJOR 2-D copies, zero/copy,
// and the "tall x small" product expressed through the dense DMMA kernel.
#include "common.cuh"
#include "nb_list.h"

namespace gsi {

int nb_for_cols(int64_t cols) {
    static const int list[] = {
#define GSI_ITEM(N) N,
        GSI_NB_LIST(GSI_ITEM)
#undef GSI_ITEM
    };
    for (int nb : list)
        if ((int64_t)8 * nb >= cols) return nb;
    return -1;
}

// stage: column-major block [cols][rb] (pitch rb) ; tall: [row][ld]
__global__ void colblock_to_tall_kernel(const double* __restrict__ stage, int64_t rb_pitch, int64_t nrows,
                                        int64_t cols, double* __restrict__ tall, int64_t ld) {
    __shared__ double tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int64_t c0 = (int64_t)blockIdx.y * 32;
    for (int cy = threadIdx.y; cy < 32; cy += blockDim.y) {
        const int64_t c = c0 + cy, r = r0 + threadIdx.x;
        tile[cy][threadIdx.x] = (c < cols && r < nrows) ? stage[c * rb_pitch + r] : 0.0;
    }
    __syncthreads();
    for (int ry = threadIdx.y; ry < 32; ry += blockDim.y) {
        const int64_t r = r0 + ry, c = c0 + threadIdx.x;
        if (r < nrows && c < cols) tall[r * ld + c] = tile[threadIdx.x][ry];
    }
}

__global__ void tall_to_colblock_kernel(const double* __restrict__ tall, int64_t ld, int64_t nrows, int64_t cols,
                                        double* __restrict__ stage, int64_t rb_pitch) {
    __shared__ double tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int64_t c0 = (int64_t)blockIdx.y * 32;
    for (int ry = threadIdx.y; ry < 32; ry += blockDim.y) {
        const int64_t r = r0 + ry, c = c0 + threadIdx.x;
        tile[ry][threadIdx.x] = (r < nrows && c < cols) ? tall[r * ld + c] : 0.0;
    }
    __syncthreads();
    for (int cy = threadIdx.y; cy < 32; cy += blockDim.y) {
        const int64_t c = c0 + cy, r = r0 + threadIdx.x;
        if (r < nrows && c < cols) stage[c * rb_pitch + r] = tile[threadIdx.x][cy];
    }
}

static const size_t kStageBytes = (size_t)96 << 20;

static int64_t stage_rows(int64_t cols) {
    int64_t rb = (int64_t)(kStageBytes / (cols * sizeof(double)));
    rb = rb / 32 * 32;
    if (rb < 32) rb = 32;
    return rb;
}

struct StageBuf {
    double* d = nullptr;
    gsi_ctx* ctx;
    explicit StageBuf(gsi_ctx* c) : ctx(c) { d = static_cast<double*>(pool_alloc(c, kStageBytes + 32 * 256 * 8)); }
    ~StageBuf() { pool_free(ctx, d, kStageBytes + 32 * 256 * 8); }
};

void tall_upload(gsi_buf* b, const double* host, int64_t ldh, int64_t row0, int64_t nrows) {
    gsi_ctx* ctx = b->ctx;
    GSI_REQUIRE(b->layout == GSI_LAYOUT_TALL, GSI_ERR_INVALID_ARGUMENT, "tall_upload: not a TALL buffer");
    GSI_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= b->rows, GSI_ERR_DIMENSION_MISMATCH,
                "tall_upload: row range outside buffer");
    GSI_REQUIRE(ldh >= nrows, GSI_ERR_INVALID_ARGUMENT, "tall_upload: leading dimension < rows");
    if (nrows == 0 || b->cols == 0) return;
    StageBuf st(ctx);
    const int64_t rb = stage_rows(b->cols);
    for (int64_t r = 0; r < nrows; r += rb) {
        const int64_t cur = (nrows - r < rb) ? nrows - r : rb;
        GSI_CUDA(cudaMemcpy2DAsync(st.d, rb * 8, host + r, ldh * 8, cur * 8, b->cols, cudaMemcpyHostToDevice,
                                   ctx->stream));
        dim3 grid((unsigned)((cur + 31) / 32), (unsigned)((b->cols + 31) / 32)), block(32, 8);
        colblock_to_tall_kernel<<<grid, block, 0, ctx->stream>>>(st.d, rb, cur, b->cols,
                                                                 b->d + (row0 + r) * b->ld, b->ld);
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx);
        // the staging area is reused by the next block
    }
    GSI_CUDA(cudaStreamSynchronize(ctx->stream));
}

void tall_download(const gsi_buf* b, double* host, int64_t ldh, int64_t row0, int64_t nrows) {
    gsi_ctx* ctx = b->ctx;
    GSI_REQUIRE(b->layout == GSI_LAYOUT_TALL, GSI_ERR_INVALID_ARGUMENT, "tall_download: not a TALL buffer");
    GSI_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= b->rows, GSI_ERR_DIMENSION_MISMATCH,
                "tall_download: row range outside buffer");
    GSI_REQUIRE(ldh >= nrows, GSI_ERR_INVALID_ARGUMENT, "tall_download: leading dimension < rows");
    if (nrows == 0 || b->cols == 0) return;
    StageBuf st(ctx);
    const int64_t rb = stage_rows(b->cols);
    for (int64_t r = 0; r < nrows; r += rb) {
        const int64_t cur = (nrows - r < rb) ? nrows - r : rb;
        dim3 grid((unsigned)((cur + 31) / 32), (unsigned)((b->cols + 31) / 32)), block(32, 8);
        tall_to_colblock_kernel<<<grid, block, 0, ctx->stream>>>(b->d + (row0 + r) * b->ld, b->ld, cur, b->cols,
                                                                 st.d, rb);
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx);
        GSI_CUDA(cudaMemcpy2DAsync(host + r, ldh * 8, st.d, rb * 8, cur * 8, b->cols, cudaMemcpyDeviceToHost,
                                   ctx->stream));
    }
    GSI_CUDA(cudaStreamSynchronize(ctx->stream));
}

void colmajor_upload(gsi_buf* b, const double* host, int64_t ldh) {
    GSI_REQUIRE(b->layout == GSI_LAYOUT_COLMAJOR, GSI_ERR_INVALID_ARGUMENT, "colmajor_upload: wrong layout");
    GSI_REQUIRE(ldh >= b->rows, GSI_ERR_INVALID_ARGUMENT, "colmajor_upload: leading dimension < rows");
    if (b->rows == 0 || b->cols == 0) return;
    GSI_CUDA(cudaMemcpy2DAsync(b->d, b->ld * 8, host, ldh * 8, b->rows * 8, b->cols, cudaMemcpyHostToDevice,
                               b->ctx->stream));
    GSI_CUDA(cudaStreamSynchronize(b->ctx->stream));
}

void colmajor_download(const gsi_buf* b, double* host, int64_t ldh) {
    GSI_REQUIRE(b->layout == GSI_LAYOUT_COLMAJOR, GSI_ERR_INVALID_ARGUMENT, "colmajor_download: wrong layout");
    GSI_REQUIRE(ldh >= b->rows, GSI_ERR_INVALID_ARGUMENT, "colmajor_download: leading dimension < rows");
    if (b->rows == 0 || b->cols == 0) return;
    GSI_CUDA(cudaMemcpy2DAsync(host, ldh * 8, b->d, b->ld * 8, b->rows * 8, b->cols, cudaMemcpyDeviceToHost,
                               b->ctx->stream));
    GSI_CUDA(cudaStreamSynchronize(b->ctx->stream));
}

void tall_zero(gsi_ctx* ctx, gsi_buf* b) {
    GSI_CUDA(cudaMemsetAsync(b->d, 0, b->bytes(), ctx->stream));
}

void tall_copy(gsi_ctx* ctx, const gsi_buf* src, gsi_buf* dst) {
    GSI_REQUIRE(src->layout == dst->layout && src->rows == dst->rows && src->cols == dst->cols &&
                    src->ld == dst->ld,
                GSI_ERR_DIMENSION_MISMATCH, "buffer copy: shape mismatch");
    GSI_CUDA(cudaMemcpyAsync(dst->d, src->d, src->bytes(), cudaMemcpyDeviceToDevice, ctx->stream));
}

// small column-major device matrix (rows x cols, ld ldm) -> TALL buffer (zero padded)
__global__ void small_cm_to_tall_kernel(const double* __restrict__ M, int64_t ldm, int64_t rows, int64_t cols,
                                        double* __restrict__ T, int64_t ld) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const int64_t r = idx % rows, c = idx / rows;
    T[r * ld + c] = M[c * ldm + r];
}

void small_cm_to_tall(gsi_ctx* ctx, const double* M, int64_t ldm, int64_t rows, int64_t cols, gsi_buf* T) {
    GSI_REQUIRE(T->layout == GSI_LAYOUT_TALL && T->rows == rows && T->cols >= cols, GSI_ERR_DIMENSION_MISMATCH,
                "small_cm_to_tall: shape mismatch");
    tall_zero(ctx, T);
    const int64_t total = rows * cols;
    if (total == 0) return;
    small_cm_to_tall_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(M, ldm, rows, cols, T->d, T->ld);
    GSI_CUDA(cudaGetLastError());
    count_launch(ctx);
}

// out[rows x lp] = Q[rows x l] * M   with M given as a TALL l x l2 buffer.  A TALL buffer
// (row-major, pitch ld) is bit-identical to the column-major matrix Q' (ld x rows), so this
// is the transposed dense DMMA GEMM on a view.
void tall_times_small(gsi_ctx* ctx, const gsi_buf* Q, const gsi_buf* Mtall, gsi_buf* out) {
    GSI_REQUIRE(Mtall->rows == Q->cols, GSI_ERR_DIMENSION_MISMATCH, "tall_times_small: inner dimension");
    GSI_REQUIRE(out->rows == Q->rows && out->cols == Mtall->cols, GSI_ERR_DIMENSION_MISMATCH,
                "tall_times_small: output shape");
    gsi_buf view;
    view.ctx = ctx; view.layout = GSI_LAYOUT_COLMAJOR;
    view.rows = Q->cols; view.cols = Q->rows; view.ld = Q->ld; view.d = Q->d; view.owns = false;
    dense_apply(ctx, &view, 1, Mtall, out, 1.0);
}

}  // namespace gsi
