"""Case list shared by oracle/make_goldens_sampling.py (which runs the compiled reference) and the parity tests of the
index-sampling utilities (RandBLAS/util.hh:459-560). Shapes follow test/test_basic_rng/test_discrete.cc."""
import numpy as np

KEYS = (42, 0, 1)
# (n, k): test_discrete.cc:169-190 (40, 17 / 34), :72-83 smoke sizes, block-boundary values of k, large n
UNIFORM_CASES = [(40, 17), (40, 34), (100, 2500), (7, 1), (1, 5), (10, 0), (2048, 1025), (2147483647, 600),
                 (1 << 40, 600), (1000003, 1024)]
CDF_COUNTER_OFFSET = 8675309                                  # test_discrete.cc:194


def uniform_cases():
    """(n, k, key, counter offset): every key at offset 0; key 42 also at the offset of test_discrete.cc:171 and
    at one that carries into the second counter limb."""
    for (n, k) in UNIFORM_CASES:
        for key in KEYS:
            for off in ((0, 3456, (1 << 32) - 2) if key == 42 else (0,)):
                yield n, k, key, off


def weight_vectors():
    """name -> (weights as float64, error_if_below or None). Cast to the scalar type under test."""
    rng = np.random.default_rng(20240917)
    N = 100
    even = np.zeros(N)
    even[::2] = 1.0 / (np.arange(0, N, 2) + 1.0)
    even[10] = 0.0                                            # test_discrete.cc:142-146
    delta = np.zeros(N)
    delta[17] = 99.0
    delta[3] = -np.finfo(np.float32).eps / 10                 # clipped without error, :154-158
    bad = np.abs(rng.standard_normal(300))
    bad[57] = -1.0                                            # below error_if_below: the reference throws at 57
    out = {
        "ones29": (np.ones(29), None),                        # :198-199
        "power100": ((1.0 / (np.arange(N) + 1.0)) ** 1.5, None),   # :122-126
        "even100": (even, None),
        "delta100": (delta, None),
        "rand10000": (np.abs(rng.standard_normal(10000)) + 1e-3, None),       # three 4096-element chunks on the device
        "rand4096": (rng.random(4096), None),
        "rand4097": (rng.random(4097) * 1e-3, None),
        "bad300": (bad, None),
        "zeros50": (np.zeros(50), None),                      # total below sqrt(n) * eps: the reference throws
        "neg_ok": (np.array([0.5, -0.25, 1.0, 2.0]), -0.5),   # explicit error_if_below
        "one": (np.array([3.0]), None),
    }
    return out


# (weights name, k) for sample_indices_iid
CDF_CASES = [("ones29", 13), ("ones29", 26), ("power100", 3000), ("even100", 3000), ("delta100", 500),
             ("rand10000", 3000), ("rand4097", 1027), ("one", 7)]
