"""-m gpu: the thread-safety contract of the C ABI (include/fastlanes_b200.h: "Thread-safe; stream-ordered; the caller may
issue from many host threads on different streams").  The reference is `Send + Sync` by construction (pure functions,
SURVEY.md §8b); here several host threads drive the device family on their own streams and the host family
concurrently, and every result must still be bit-exact against the oracle."""
import threading

import numpy as np
import pytest

from gpu_util import dev_empty, rand_bytes, to_dev, to_host

pytestmark = pytest.mark.gpu


def test_concurrent_streams_and_host_calls(oracle):
    import torch

    import fastlanes_b200 as fl

    n_threads, n_blocks, rounds = 6, 257, 8
    errors = []

    def device_worker(tid):
        try:
            torch.cuda.set_device(0)
            rng = np.random.default_rng(5000 + tid)
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                for r in range(rounds):
                    w = (tid * 5 + r * 3) % 33
                    packed = rand_bytes(rng, n_blocks * 128 * w, 32)
                    out = dev_empty(n_blocks * 1024, 32)
                    fl.BitPacking.unpack(w, to_dev(packed), out)
                    back = dev_empty(n_blocks * 32 * w, 32)
                    fl.BitPacking.pack(w, out, back)
                    stream.synchronize()
                    assert np.array_equal(to_host(out, 32), oracle.unpack(packed, w, n_blocks=n_blocks)), (tid, r, w)
                    assert np.array_equal(to_host(back, 32), packed), (tid, r, w, "pack")
        except Exception as e:  # noqa: BLE001
            errors.append(("device", tid, repr(e)))

    def host_worker(tid):
        try:
            torch.cuda.set_device(0)
            rng = np.random.default_rng(6000 + tid)
            for r in range(rounds):
                tb = (8, 16, 32, 64)[(tid + r) % 4]
                w = (tid * 7 + r) % (tb + 1)
                packed = rand_bytes(rng, n_blocks * 128 * w, tb)
                out = np.empty(n_blocks * 1024, dtype=packed.dtype)
                fl.BitPacking.unpack(w, packed, out)
                assert np.array_equal(out, oracle.unpack(packed, w, n_blocks=n_blocks)), (tid, r, tb, w)
                bitmap = np.empty(n_blocks * 128, dtype=np.uint8)
                fl.Scan.filter_range(w, packed, 0, 1, 1 << max(0, w - 1), bitmap)
                sel = (out >= 1) & (out <= (1 << max(0, w - 1)))
                assert np.array_equal(bitmap, np.packbits(sel, bitorder="little")), (tid, r, tb, w, "filter")
        except Exception as e:  # noqa: BLE001
            errors.append(("host", tid, repr(e)))

    threads = [threading.Thread(target=device_worker, args=(t,)) for t in range(n_threads)]
    threads += [threading.Thread(target=host_worker, args=(t,)) for t in range(3)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors[:3]
