// randblas_b200 -- aggregate header of the drop-in layer (the counterpart of the reference's RandBLAS.hh:33-41
// for the sketching hot path). Header-only host code; link librandblas_b200.so.
#pragma once
#include "RandBLAS/base.hh"
#include "RandBLAS/dense_skops.hh"
#include "RandBLAS/sparse_skops.hh"
#include "RandBLAS/sparse_data.hh"
#include "RandBLAS/sketch.hh"
#include "RandBLAS/util.hh"
#include "RandBLAS/multi_gpu.hh"
