// Multi-GPU ghost-row summation (placeholder).
#include "gtk_internal.h"
void gtk_comm_release(gtk_ctx* ctx) { (void)ctx; }
extern "C" {
int32_t gtk_comm_unique_id(void* id128) { (void)id128; return GTK_ERR_NCCL; }
int32_t gtk_comm_init(gtk_ctx* ctx, int32_t rank, int32_t n_ranks, const void* id128) {
  (void)rank; (void)n_ranks; (void)id128; if (!ctx) return GTK_ERR_INVALID; GTK_FAIL(GTK_ERR_NCCL, "not built yet");
}
int32_t gtk_comm_setup_ghost_rows(gtk_ctx* ctx, int64_t lo, int64_t hi) {
  (void)lo; (void)hi; if (!ctx) return GTK_ERR_INVALID; GTK_FAIL(GTK_ERR_NCCL, "not built yet");
}
int64_t gtk_comm_ghost_info(const gtk_ctx* ctx, int32_t key) { (void)ctx; (void)key; return 0; }
int32_t gtk_comm_sum_ghost_rows(gtk_ctx* ctx) { if (!ctx) return GTK_ERR_INVALID; GTK_FAIL(GTK_ERR_NCCL, "not built yet"); }
}
