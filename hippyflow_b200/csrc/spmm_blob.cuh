// Per-cluster record ("blob") of the packed CSR matrix read by the panel DMMA SpMM (csr_spmm_dmma_kernel, spmm_dmma.cu; the
// retired variants under tools/experiments/spmm_variants/ used it too).  Written on the host by hfb_csr_pack_clusters:
//     int32 header[4] = {nrow, ncol, nent, 0}
//     int32 rowoff[max_rows + 1]   entry offsets of the cluster's rows (relative to the cluster's first entry)
//     int32 outrow[max_rows]       global row index of each cluster row (where its result goes in C)
//     int32 cols[max_cols]         the distinct global columns (= rows of B) the cluster touches
//     {double v; int32 l; int32 r} entries[nent]   value, cluster-LOCAL column index, cluster-LOCAL row index
#pragma once

namespace hfb {

struct SpmmBlobLayout {
    int off_rowoff, off_outrow, off_cols, off_ent, stride;  // bytes
};

static inline int round_up(int x, int a) { return (x + a - 1) / a * a; }

static SpmmBlobLayout blob_layout(int max_rows, int max_cols, int max_entries) {
    SpmmBlobLayout L;
    L.off_rowoff = 16;
    L.off_outrow = L.off_rowoff + 4 * round_up(max_rows + 1, 4);
    L.off_cols = L.off_outrow + 4 * round_up(max_rows, 4);
    L.off_ent = L.off_cols + 4 * round_up(max_cols, 4);
    L.stride = round_up(L.off_ent + 16 * max_entries, 128);
    return L;
}

}  // namespace hfb
