// Instantiates the pair kernels for one scalar core (see lec_pairs_impl.cuh).
#include "lec_pairs_impl.cuh"
LEC_DEFINE_CORE_TU(lec::CORE_EUC32, euc32)
