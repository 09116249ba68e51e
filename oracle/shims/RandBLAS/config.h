// TEST INFRASTRUCTURE (oracle shim) -- hand-written stand-in for the CMake-generated
// RandBLAS/config.h (template: RandBLAS/config.h.in:44-54 in the reference).
#pragma once
#define RandBLAS_FULL_VERSION "oracle-shim"
#define RandBLAS_VERSION_MAJOR 1
#define RandBLAS_VERSION_MINOR 1
#define RandBLAS_VERSION_PATCH 0
#define RandBLAS_VERSION_DEVEL 0
#define RandBLAS_HAS_OpenMP
