"""Output path of the per-view driver: the reference's PFM / confidence writers
(``datasets/data_io.py:45-72``, ``test.py:268-293``) behind a background writer, so that the consumer of
``StreamedCascade.run`` hands a finished view to disk without stalling the GPU pipeline.

* ``save_pfm(filename, image, scale=1)`` / ``read_pfm(filename)`` — byte-compatible with the reference's.
* ``AsyncResultWriter`` — bounded queue + worker threads; ``submit(depth_path, depth, conf_path, confidence)`` copies
  nothing (the arrays handed over must not be reused by the caller: ``StreamedCascade`` yields fresh host arrays or
  ring slots — pass ``copy=True`` for ring slots), ``close()`` drains and re-raises the first I/O error.
"""
import os
import queue
import re
import sys
import threading

import numpy as np


def save_pfm(filename, image, scale=1):
    """datasets/data_io.py:45-72: 'Pf' / 'PF' header, width height, negative scale for little endian, rows bottom-up."""
    image = np.flipud(image)
    if image.dtype.name != "float32":
        raise Exception("Image dtype must be float32.")
    if len(image.shape) == 3 and image.shape[2] == 3:
        color = True
    elif len(image.shape) == 2 or len(image.shape) == 3 and image.shape[2] == 1:
        color = False
    else:
        raise Exception("Image must have H x W x 3, H x W x 1 or H x W dimensions.")
    endian = image.dtype.byteorder
    if endian == "<" or endian == "=" and sys.byteorder == "little":
        scale = -scale
    with open(filename, "wb") as f:
        f.write(("PF\n" if color else "Pf\n").encode("utf-8"))
        f.write("{} {}\n".format(image.shape[1], image.shape[0]).encode("utf-8"))
        f.write(("%f\n" % scale).encode("utf-8"))
        image.tofile(f)


def read_pfm(filename):
    """datasets/data_io.py:7-42 -> (data, scale)."""
    with open(filename, "rb") as f:
        header = f.readline().decode("utf-8").rstrip()
        if header not in ("PF", "Pf"):
            raise Exception("Not a PFM file.")
        color = header == "PF"
        dim_match = re.match(r"^(\d+)\s(\d+)\s$", f.readline().decode("utf-8"))
        if not dim_match:
            raise Exception("Malformed PFM header.")
        width, height = map(int, dim_match.groups())
        scale = float(f.readline().rstrip())
        endian = "<" if scale < 0 else ">"
        data = np.fromfile(f, endian + "f")
    shape = (height, width, 3) if color else (height, width)
    return np.flipud(np.reshape(data, shape)), abs(scale)


class AsyncResultWriter:
    """Writes ``depth_est`` PFMs and ``confidence`` .npy files (test.py:268-293) on background threads."""

    def __init__(self, workers=2, max_pending=8):
        self._q = queue.Queue(maxsize=max_pending)
        self._error = None
        self._threads = [threading.Thread(target=self._run, daemon=True) for _ in range(workers)]
        for t in self._threads:
            t.start()

    def _run(self):
        while True:
            job = self._q.get()
            try:
                if job is None:
                    return
                depth_path, depth, conf_path, conf = job
                if self._error is None:
                    if depth_path is not None:
                        os.makedirs(os.path.dirname(depth_path) or ".", exist_ok=True)
                        save_pfm(depth_path, depth)
                    if conf_path is not None:
                        os.makedirs(os.path.dirname(conf_path) or ".", exist_ok=True)
                        np.save(conf_path, conf)
            except Exception as exc:                    # surfaced by close() / the next submit()
                self._error = exc
            finally:
                self._q.task_done()

    def submit(self, depth_path, depth, conf_path=None, confidence=None, copy=True):
        if self._error is not None:
            raise self._error
        depth = np.asarray(depth, dtype=np.float32)
        if copy:
            depth = depth.copy()
            confidence = None if confidence is None else np.array(confidence, copy=True)
        self._q.put((depth_path, depth, conf_path, confidence))

    def close(self):
        self._q.join()
        for _ in self._threads:
            self._q.put(None)
        for t in self._threads:
            t.join()
        if self._error is not None:
            raise self._error

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
