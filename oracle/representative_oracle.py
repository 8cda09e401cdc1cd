"""CPU restatement of the reference's representative-descriptor choice.  TEST INFRASTRUCTURE ONLY.

Follows reference slam/nodes.py:146-153 (MapPoint.add_observation) literally:
    dists = np.zeros((n, n)); for i1 < i2: dists[i2, i1] = dists[i1, i2] = dbow.distance(f1, f2)
    self.feat = self.obs[np.argmin(np.median(dists, axis=0))][1]
dbow.distance is pyDBoW3 Vocabulary.distance -> DBoW3 hamming_distance (docker/pydbow3/src/dbow3.cpp:110-112;
the DBoW3 fork itself is not vendored, docker/install_dbow3.sh:7): popcount(f1 XOR f2) over the 32 bytes.
One deliberate difference: the reference stores the distances in a uint8 array, so a distance of exactly
256 (all bits differ) would not fit; here distances are kept as integers.  Parity is otherwise pinned by
this literal loop (the reference has no test for this step).
"""
import numpy as np

_POP = np.array([bin(i).count("1") for i in range(256)], np.int64)


def hamming(f1, f2) -> int:
    return int(_POP[np.bitwise_xor(np.asarray(f1, np.uint8), np.asarray(f2, np.uint8))].sum())


def select(obs) -> int:
    """obs: sequence of n >= 1 descriptors uint8[32]; returns the index the reference would keep."""
    n = len(obs)
    dists = np.zeros((n, n), np.int64)
    for i1 in range(n):
        for i2 in range(n):
            if i1 < i2:
                dists[i2, i1] = dists[i1, i2] = hamming(obs[i1], obs[i2])
    return int(np.argmin(np.median(dists, axis=0)))


def select_batch(obs, counts) -> np.ndarray:
    return np.array([select(obs[p, :c]) if c > 0 else -1 for p, c in enumerate(np.asarray(counts))], np.int32)
