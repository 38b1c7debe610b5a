// Error reporting and device queries of the C ABI (include/idv.h).
#include <stdarg.h>
#include <string.h>

#include "idv_common.cuh"

namespace idv {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static int g_lstm_ncols = 0;
static int g_dynamic_tiles = 0;
static int g_gemm_pairs = 1;
static int g_wave_pairs = 1;
static int g_tile_order = 1;
static int g_tma_store = 1;
static int g_lstm_interleave = 1;
static int g_lstm_sync_mode = 0;
static int g_launch_pdl = 1;
static int g_splitk = 1;
static int g_lstm_cluster_alt = 0;
static int g_lstm_chunk_sync = 1;
static int g_lstm_tma_publish = -1;
int option_splitk() { return g_splitk; }
int option_lstm_cluster_alt() { return g_lstm_cluster_alt; }
int option_lstm_chunk_sync() { return g_lstm_chunk_sync; }
int option_lstm_tma_publish() { return g_lstm_tma_publish; }
int option_launch_pdl() { return g_launch_pdl; }
int option_lstm_sync_mode() { return g_lstm_sync_mode; }
int option_lstm_interleave() { return g_lstm_interleave; }
int option_tma_store() { return g_tma_store; }
int option_tile_order() { return g_tile_order; }
int option_lstm_wave_pairs() { return g_wave_pairs; }
int option_gemm_pairs() { return g_gemm_pairs; }
int option_lstm_ncols() { return g_lstm_ncols; }
int option_dynamic_tiles() { return g_dynamic_tiles; }
}  // namespace idv

extern "C" int idv_set_option(const char* name, int value) {
  using namespace idv;
  IDV_CHECK_ARG(name, "idv_set_option: null name");
  if (strcmp(name, "lstm_ncols") == 0) {
    IDV_CHECK_ARG(value == 0 || value == 32 || value == 48 || value == 64, "idv_set_option: lstm_ncols must be 0, 32, 48 or 64");
    g_lstm_ncols = value;
    return IDV_OK;
  }
  if (strcmp(name, "gemm_dynamic_tiles") == 0) {
    g_dynamic_tiles = value != 0;
    return IDV_OK;
  }
  if (strcmp(name, "lstm_wave_cta_pairs") == 0) {
    g_wave_pairs = value != 0;
    return IDV_OK;
  }
  if (strcmp(name, "gemm_tile_order") == 0) {
    IDV_CHECK_ARG(value == 0 || value == 1, "idv_set_option: gemm_tile_order must be 0 (unit-major) or 1 (row-major)");
    g_tile_order = value;
    return IDV_OK;
  }
  if (strcmp(name, "lstm_sync_mode") == 0) {
    IDV_CHECK_ARG(value >= 0 && value <= 2, "idv_set_option: lstm_sync_mode must be 0 (fence + atomic, acquire polls), 1 (red.release, relaxed polls + fence) or 2 (red.release, acquire polls)");
    g_lstm_sync_mode = value;
    return IDV_OK;
  }
  if (strcmp(name, "gemm_splitk") == 0) {
    g_splitk = value != 0;
    return IDV_OK;
  }
  if (strcmp(name, "lstm_tma_publish") == 0) {
    IDV_CHECK_ARG(value >= -1 && value <= 2, "idv_set_option: lstm_tma_publish must be -1 (auto), 0 (direct stores), 1 (tensor store) or 2 (staged 16-byte stores)");
    g_lstm_tma_publish = value;
    return IDV_OK;
  }
  if (strcmp(name, "lstm_chunk_sync") == 0) {
    g_lstm_chunk_sync = value != 0;
    return IDV_OK;
  }
  if (strcmp(name, "lstm_cluster_alt") == 0) {
    IDV_CHECK_ARG(value == 0 || value == 1, "idv_set_option: lstm_cluster_alt must be 0 (largest CTAs) or 1 (second choice)");
    g_lstm_cluster_alt = value;
    return IDV_OK;
  }
  if (strcmp(name, "launch_pdl") == 0) {
    g_launch_pdl = value != 0;
    return IDV_OK;
  }
  if (strcmp(name, "lstm_interleave") == 0) {
    g_lstm_interleave = value != 0;
    return IDV_OK;
  }
  if (strcmp(name, "gemm_tma_store") == 0) {
    g_tma_store = value != 0;
    return IDV_OK;
  }
  if (strcmp(name, "gemm_cta_pairs") == 0) {
    g_gemm_pairs = value != 0;
    return IDV_OK;
  }
  set_error("idv_set_option: unknown option %s", name);
  return IDV_E_ARG;
}

extern "C" int idv_abi_version(void) { return IDV_ABI_VERSION; }
extern "C" const char* idv_last_error(void) { return idv::g_err; }
extern "C" int idv_device_sm_count(int* out) {
  using namespace idv;
  IDV_CHECK_ARG(out, "idv_device_sm_count: null pointer");
  int dev = 0;
  IDV_CUDA(cudaGetDevice(&dev));
  IDV_CUDA(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
  return IDV_OK;
}
