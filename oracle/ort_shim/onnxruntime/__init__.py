"""ORACLE shim (test infrastructure): a module *named* onnxruntime so that the
unmodified reference package (/root/reference/Skps) imports and runs here, where
the real onnxruntime is not installed.  Executes the same .onnx graphs with
oracle.onnx_exec (PyTorch CPU fp32).  Mirrors only the API the reference touches
(/root/reference/Skps/core/api/onnx_model_base.py:14,23)."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_root = os.path.abspath(os.path.join(_here, "..", "..", ".."))
if _root not in sys.path:
    sys.path.append(_root)   # appended: the reference tree must win for Skps, core, logger

from oracle.onnx_exec import Session as _Session  # noqa: E402


class _Input:
    def __init__(self, name):
        self.name = name


class InferenceSession:
    def __init__(self, path, providers=None, **kw):
        self._s = _Session(path)

    def get_inputs(self):
        return [_Input(self._s.input_name)]

    def run(self, output_names, feed):
        (x,) = feed.values()
        return self._s.run(x)
