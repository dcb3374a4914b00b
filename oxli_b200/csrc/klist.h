// klist.h -- the k values with a compile-time specialised consume kernel.
// One translation unit per k (consume_inst.cu, -DOXG_INST_K=k) so that they compile in
// parallel; oxli_b200/_build.py reads this list, capi.cu dispatches over it.  Every other
// k in 1..255 runs consume_generic_kernel.  The partitioned and sharded pipelines exist for every listed k.
#pragma once
#define OXG_FOR_EACH_K(X) \
    X(15) X(17) X(19) X(20) X(21) X(23) X(24) X(25) X(27) X(29) X(31) X(32) \
    X(33) X(35) X(37) X(39) X(41) X(43) X(45) X(47) X(49) X(51) X(53) X(55) X(57) X(59) X(61) X(63)
