// NVTX ranges around the C-ABI entry points (host side), one per kernel family: csm.create (weight packing),
// csm.generate_frame > csm.prefill / csm.decode.mega / csm.decode.graph, mimi.encode / mimi.decode / mimi.decode_stream,
// csm.post.*.  nvtx3 is header-only and resolves its injection library lazily: without a profiler attached a push / pop
// is a call through a null-checked function pointer (nanoseconds against a 3 ms frame).  `ncu --nvtx --nvtx-include
// "csm.decode.mega/"` restricts a capture to one family (profiles/README.md).
#pragma once
#include <nvtx3/nvToolsExt.h>

struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
