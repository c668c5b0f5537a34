// Layout plumbing for device buffers: host column-major <-> device TALL (row-major,
// pitch ld) conversion through a bounded staging area, COLMAJOR 2-D copies, zero/copy,
// and the "tall x small" product expressed through the dense DMMA kernel.
#include "common.cuh"
#include "nb_list.h"
#include <cstring>
#include <thread>
#include <vector>

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

// Host (pageable or pinned, pitched) -> device (pitched) 2-D copy.  A pageable source goes through two pinned
// bounce buffers of the context (host memcpy of the next slice overlaps the DMA of the previous one): the
// driver's own staging of a pageable cudaMemcpy2DAsync ran at ~2.3 GB/s here (82 MB rga batch: 35 ms).
static const size_t kPinBytes = (size_t)16 << 20;

// Pageable source -> bounce buffer.  One core copies at 3-10 GB/s, less than half of what the DMA engine
// moves, so slices of a few MB and more are split over a handful of threads.
static void host_copy_rows(char* dst, const char* src, size_t nr, size_t width, size_t spitch) {
    const size_t total = nr * width;
    unsigned hw = std::thread::hardware_concurrency();
    size_t nt = hw >= 16 ? 8 : hw >= 8 ? 4 : hw >= 4 ? 2 : 1;
    if (total < ((size_t)4 << 20)) nt = 1;
    auto copy_range = [=](size_t b0, size_t b1) {             // byte range of the packed destination
        if (spitch == width) { memcpy(dst + b0, src + b0, b1 - b0); return; }
        size_t b = b0;
        while (b < b1) {
            const size_t row = b / width, off = b - row * width;
            const size_t len = (width - off < b1 - b) ? width - off : b1 - b;
            memcpy(dst + b, src + row * spitch + off, len);
            b += len;
        }
    };
    if (nt <= 1) { copy_range(0, total); return; }
    const size_t chunk = ((total + nt - 1) / nt + 4095) / 4096 * 4096;
    std::vector<std::thread> th;
    th.reserve(nt);
    size_t started_to = chunk < total ? chunk : total;          // bytes [chunk, started_to) are with helper threads
    for (size_t t = 1; t < nt; ++t) {
        const size_t b0 = t * chunk, b1 = (b0 + chunk < total) ? b0 + chunk : total;
        if (b0 >= total) break;
        try {
            th.emplace_back(copy_range, b0, b1);
        } catch (...) {
            break;                                              // no more threads to be had: this thread copies the rest
        }
        started_to = b1;
    }
    copy_range(0, chunk < total ? chunk : total);
    if (started_to < total) copy_range(started_to, total);
    for (auto& x : th) x.join();
}

static void h2d_2d(gsi_ctx* ctx, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height) {
    if (width == 0 || height == 0) return;
    cudaPointerAttributes at;
    const bool pinned_src = cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned_src || width > kPinBytes) {
        GSI_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyHostToDevice, ctx->stream));
        return;
    }
    if (!ctx->pin[0]) {
        for (int i = 0; i < 2; ++i) {
            GSI_CUDA(cudaHostAlloc(&ctx->pin[i], kPinBytes, cudaHostAllocDefault));
            GSI_CUDA(cudaEventCreateWithFlags(&ctx->pin_ev[i], cudaEventDisableTiming));
        }
    }
    const size_t rows_per = kPinBytes / width;
    int b = 0;
    for (size_t r = 0; r < height; r += rows_per, b ^= 1) {
        const size_t nr = (height - r < rows_per) ? height - r : rows_per;
        GSI_CUDA(cudaEventSynchronize(ctx->pin_ev[b]));            // the DMA that last read this bounce buffer is done
        char* pb = static_cast<char*>(ctx->pin[b]);
        const char* sp = static_cast<const char*>(src) + r * spitch;
        host_copy_rows(pb, sp, nr, width, spitch);
        GSI_CUDA(cudaMemcpy2DAsync(static_cast<char*>(dst) + r * dpitch, dpitch, pb, width, width, nr, cudaMemcpyHostToDevice,
                                   ctx->stream));
        GSI_CUDA(cudaEventRecord(ctx->pin_ev[b], ctx->stream));
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
        h2d_2d(ctx, st.d, rb * 8, host + r, ldh * 8, cur * 8, b->cols);
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
    h2d_2d(b->ctx, b->d, b->ld * 8, host, ldh * 8, b->rows * 8, b->cols);
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

__global__ void tall_cols_copy_kernel(const double* __restrict__ src, int64_t lds, double* __restrict__ dst, int64_t ldd,
                                      int64_t rows, int w) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = idx / w;
    const int c = (int)(idx - r * w);
    if (r < rows) dst[r * ldd + c] = src[r * lds + c];
}

void tall_cols_copy(gsi_ctx* ctx, const double* src, int64_t lds, double* dst, int64_t ldd, int64_t rows, int64_t w) {
    if (rows <= 0 || w <= 0) return;
    const int64_t total = rows * w;
    tall_cols_copy_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(src, lds, dst, ldd, rows, (int)w);
    GSI_CUDA(cudaGetLastError());
    count_launch(ctx);
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
