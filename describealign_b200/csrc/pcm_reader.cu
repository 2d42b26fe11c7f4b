// Decode hand-off (reference describealign.py:149-157).
//
// The reference lets ffmpeg write a whole track as s16le to a pipe, collects it in a Python bytes object
// (`capture_stdout`), converts it to float16 on the host and only then starts computing.  Here a reader thread
// inside the library takes the decoder's pipe itself: it read()s into two page-locked chunks in turn and copies each
// chunk to the device on its own stream while the decoder is still producing, so that when the pipe reaches EOF the
// track is already in HBM as int16 - the feature kernel converts to float16 on the fly (features.cu) - and the decode
// of the next pair overlaps the alignment of the current one.  No Python object ever holds the samples.
#include <errno.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <string>
#include <thread>

#include "common.cuh"

struct dab_pcm_reader {
  dab_ctx *ctx = nullptr;
  int fd = -1;
  std::thread th;
  cudaStream_t stream = nullptr;
  void *chunk[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  size_t chunk_bytes = 0;
  void *dev = nullptr;
  size_t cap = 0;
  std::atomic<int64_t> bytes{0};      // bytes handed to the copy engine so far
  std::atomic<int> finished{0};
  int rc = DAB_OK;
  std::string err;
};

namespace {

constexpr size_t CHUNK = 8u << 20;    // 8 MiB per page-locked chunk: ~0.15 ms of PCIe time, far above the copy set-up cost

int reader_fail(dab_pcm_reader *r, const std::string &msg) {
  r->err = msg;
  r->rc = DAB_E_CUDA;
  return DAB_E_CUDA;
}

int grow(dab_pcm_reader *r, size_t need) {
  if (need <= r->cap) return DAB_OK;
  size_t ncap = r->cap ? r->cap : (64u << 20);
  while (ncap < need) ncap *= 2;
  void *nd = nullptr;
  if (cudaMallocAsync(&nd, ncap, r->stream) != cudaSuccess) return reader_fail(r, "pcm reader: out of device memory");
  if (r->dev) {
    const size_t have = (size_t)r->bytes.load();
    if (have && cudaMemcpyAsync(nd, r->dev, have, cudaMemcpyDeviceToDevice, r->stream) != cudaSuccess)
      return reader_fail(r, "pcm reader: device copy failed");
    cudaFreeAsync(r->dev, r->stream);
  }
  r->dev = nd;
  r->cap = ncap;
  return DAB_OK;
}

void reader_main(dab_pcm_reader *r) {
  if (cudaSetDevice(r->ctx->device) != cudaSuccess) { reader_fail(r, "pcm reader: cudaSetDevice failed"); r->finished = 1; return; }
  int which = 0;
  bool used[2] = {false, false};
  for (;;) {
    // the chunk may still be the source of a copy in flight
    if (used[which] && cudaEventSynchronize(r->done[which]) != cudaSuccess) { reader_fail(r, "pcm reader: event wait failed"); break; }
    size_t got = 0;
    bool eof = false;
    while (got < r->chunk_bytes) {
      const ssize_t n = ::read(r->fd, static_cast<char *>(r->chunk[which]) + got, r->chunk_bytes - got);
      if (n > 0) { got += (size_t)n; continue; }
      if (n == 0) { eof = true; break; }
      if (errno == EINTR) continue;
      reader_fail(r, std::string("pcm reader: read() failed: ") + strerror(errno));
      eof = true;
      break;
    }
    if (r->rc != DAB_OK) break;
    if (got) {
      const size_t at = (size_t)r->bytes.load();
      if (grow(r, at + got) != DAB_OK) break;
      if (cudaMemcpyAsync(static_cast<char *>(r->dev) + at, r->chunk[which], got, cudaMemcpyHostToDevice, r->stream) != cudaSuccess ||
          cudaEventRecord(r->done[which], r->stream) != cudaSuccess) {
        reader_fail(r, "pcm reader: host-to-device copy failed");
        break;
      }
      used[which] = true;
      r->bytes.store((int64_t)(at + got));
      which ^= 1;
    }
    if (eof) break;
  }
  if (cudaStreamSynchronize(r->stream) != cudaSuccess && r->rc == DAB_OK) reader_fail(r, "pcm reader: stream sync failed");
  r->finished = 1;
}

}  // namespace

extern "C" {

int dab_pcm_reader_open(dab_ctx *ctx, int fd, int64_t expected_bytes, dab_pcm_reader **out) {
  if (!ctx || fd < 0 || !out) return DAB_E_ARG;
  DAB_CUDA(cudaSetDevice(ctx->device));
  dab_pcm_reader *r = new dab_pcm_reader();
  r->ctx = ctx;
  r->fd = fd;
  r->chunk_bytes = CHUNK;
  cudaError_t e = cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking);
  for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
    r->chunk[k] = dab_alloc_pinned(CHUNK);
    if (!r->chunk[k]) e = cudaErrorMemoryAllocation;
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->done[k], cudaEventDisableTiming);
  }
  if (e == cudaSuccess && expected_bytes > 0) {
    // the caller knows the duration (ffprobe): one allocation, no growth
    if (grow(r, (size_t)expected_bytes + 16) != DAB_OK) e = cudaErrorMemoryAllocation;
  }
  if (e != cudaSuccess) {
    dab_set_err(ctx, std::string("dab_pcm_reader_open: ") + cudaGetErrorString(e));
    dab_pcm_reader_close(r);
    return DAB_E_CUDA;
  }
  r->th = std::thread(reader_main, r);
  *out = r;
  return DAB_OK;
}

int64_t dab_pcm_reader_progress(dab_pcm_reader *r) { return r ? r->bytes.load() : 0; }

// Blocks until the pipe reached EOF and every chunk is on the device.  *device_pcm stays valid until
// dab_pcm_reader_close; hand it to dab_pair_set_pcm(..., on_device = 1).
int dab_pcm_reader_wait(dab_pcm_reader *r, void **device_pcm, int64_t *bytes) {
  if (!r || !device_pcm || !bytes) return DAB_E_ARG;
  if (r->th.joinable()) r->th.join();
  if (r->rc != DAB_OK) {
    dab_set_err(r->ctx, r->err);
    return r->rc;
  }
  *device_pcm = r->dev;
  *bytes = r->bytes.load();
  return DAB_OK;
}

// The decoded samples back on the host (int16, interleaved): --stretch_audio needs them there afterwards
// (describealign.py:1142-1150).  dst: at least `bytes` bytes, bytes <= what dab_pcm_reader_wait reported.
int dab_pcm_reader_copy_to_host(dab_pcm_reader *r, void *dst, int64_t bytes) {
  if (!r || !dst || bytes < 0) return DAB_E_ARG;
  dab_ctx *ctx = r->ctx;
  if (r->th.joinable()) r->th.join();
  if (r->rc != DAB_OK || bytes > r->bytes.load()) return DAB_E_ARG;
  if (bytes == 0) return DAB_OK;
  DAB_CUDA(cudaSetDevice(ctx->device));
  DAB_CUDA(cudaMemcpyAsync(dst, r->dev, (size_t)bytes, cudaMemcpyDeviceToHost, r->stream));
  DAB_CUDA(cudaStreamSynchronize(r->stream));
  return DAB_OK;
}

void dab_pcm_reader_close(dab_pcm_reader *r) {
  if (!r) return;
  if (r->th.joinable()) r->th.join();
  cudaSetDevice(r->ctx->device);
  if (r->stream) cudaStreamSynchronize(r->stream);
  if (r->dev) cudaFreeAsync(r->dev, r->stream);
  for (int k = 0; k < 2; ++k) {
    if (r->done[k]) cudaEventDestroy(r->done[k]);
    if (r->chunk[k]) dab_free_pinned(r->chunk[k]);
  }
  if (r->stream) { cudaStreamSynchronize(r->stream); cudaStreamDestroy(r->stream); }
  delete r;
}

}  // extern "C"
