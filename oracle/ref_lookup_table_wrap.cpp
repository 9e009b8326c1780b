// ref_lookup_table_wrap.cpp -- thin C wrapper around the REFERENCE's own LookupTable
// (/root/reference/scene_flow_clusterer/{include/lookup_table.h,src/lookup_table.cpp}), compiled from the
// sources where they lie by oracle/Makefile into oracle/_ref/.  TEST INFRASTRUCTURE ONLY: it is the one piece
// of the reference that builds without ROS/PCL/Eigen/OpenCV and it pins the oracle's EquivalenceTable.
// No reference source is copied into this repository; only the header is #included at build time.
#include <cstddef>
#include <cstdint>
#include <stdexcept>

#include "lookup_table.h"

extern "C" int ref_lookup_table_run(int size, const int32_t* ops, int n_ops, int32_t* out, int max_out) {
  LookupTable table(static_cast<size_t>(size));
  table.reset();
  int n_out = 0;
  try {
    for (int i = 0; i < n_ops; i++) {
      const int op = ops[3 * i], a = ops[3 * i + 1], b = ops[3 * i + 2];
      int r;
      if (op == 0)
        r = table.addLabel();
      else if (op == 1) {
        table.link(a, b);
        continue;
      } else
        r = table.lookup(a);
      if (n_out < max_out) out[n_out] = r;
      n_out++;
    }
  } catch (const std::out_of_range&) {
    return -1;
  }
  return n_out;
}
