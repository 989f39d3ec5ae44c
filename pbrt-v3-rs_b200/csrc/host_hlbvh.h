// Internal (not part of the C ABI): the host half of the GPU HLBVH builder.
#pragma once
#include <stdint.h>

#include "../../include/b200pt.h"

// build_upper_sah (accelerators/src/bvh/hlbvh.rs:353-449) fused with flatten_bvh_tree (mod.rs:126-153) over the treelet
// roots: final index of every treelet's block and the interior nodes above them.  upper_nodes_out / upper_index_out need
// room for n_treelets - 1 entries.  Defined in host_hlbvh.cpp.
extern "C" int b200pt_hlbvh_upper_layout(const b200pt_bvh_node* treelet_roots, const uint32_t* treelet_n_nodes, int64_t n_treelets,
                                         b200pt_bvh_node* upper_nodes_out, int64_t* upper_index_out, int64_t* n_upper_out, int64_t* treelet_base_out,
                                         int64_t* n_nodes_out);
