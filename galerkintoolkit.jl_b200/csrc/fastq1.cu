// Structured Q1 fast path (placeholder until the tile-plan kernels land): never claims a call.
#include "gtk_internal.h"
int32_t gtk_fastq1_try(gtk_ctx* ctx, int mform, const gtk_form_params* pm, int vform,
                       const gtk_form_params* pv, bool* handled) {
  (void)ctx; (void)mform; (void)pm; (void)vform; (void)pv;
  *handled = false;
  return GTK_OK;
}
void gtk_fastq1_release(gtk_ctx* ctx) { (void)ctx; }
