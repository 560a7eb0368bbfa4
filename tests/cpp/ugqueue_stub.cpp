// Preloaded in front of libugeval.so when oracle/_ref/ref_search_queue (the reference's engine with
// unrealgo_b200/host/patches/uctsearch_leaf_queue.patch) runs without a GPU: the evaluator entry points of
// ugeval_stub.c plus the stub leaf queue of stub_queue.h.  ug_pack_planes is NOT stubbed — it is host code and the
// product's own runs.  Test infrastructure only.
extern "C" {
#include "ugeval_stub.c"
}
#include "stub_queue.h"
