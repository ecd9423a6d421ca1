"""Independent NumPy restatement of the scoring function (anti-diagonal vectorised), used to cross-check
the C oracle.  Follows smith_waterman.cl:5-7 (constants) and :114-125 (recurrence); the end cell is the
first maximum in row-major order (np.argmax on the row-major matrix)."""
import numpy as np


def sw_matrix(a: bytes, b: bytes) -> np.ndarray:
    a = np.frombuffer(bytes(a), dtype=np.uint8)
    b = np.frombuffer(bytes(b), dtype=np.uint8)
    n, m = a.size, b.size
    H = np.zeros((n + 1, m + 1), dtype=np.int64)
    if n == 0 or m == 0:
        return H[1:, 1:]
    S = np.where(a[:, None] == b[None, :], 2, -1)
    for d in range(n + m - 1):               # anti-diagonal i + j = d
        i0, i1 = max(0, d - m + 1), min(n - 1, d)
        i = np.arange(i0, i1 + 1)
        j = d - i
        diag = H[i, j] + S[i, j]
        up = H[i, j + 1] - 2
        left = H[i + 1, j] - 2
        H[i + 1, j + 1] = np.maximum(np.maximum(diag, up), np.maximum(left, 0))
    return H[1:, 1:]


def sw_linear(a: bytes, b: bytes):
    H = sw_matrix(a, b)
    if H.size == 0 or H.max() == 0:
        return 0, -1, -1
    k = int(np.argmax(H))                    # first maximum, row-major
    return int(H.flat[k]), k // H.shape[1], k % H.shape[1]


def last_row_max(a: bytes, b: bytes) -> int:
    H = sw_matrix(a, b)
    return int(H[-1].max()) if H.size else 0
