"""`torch.ops.geobipy_b200.*`: the three operators of SURVEY.md section 8(b) registered with PyTorch's dispatcher.

    torch.ops.geobipy_b200.fdem_forward(system, nlayers, sigma, thickness, altitude, precision) -> pred [B, C]
    torch.ops.geobipy_b200.fdem_sensitivity(system, nlayers, sigma, thickness, altitude, precision) -> (pred [B, C], J [B, C, L])
    torch.ops.geobipy_b200.rjmcmc_run(system, options, data, altitude, seed, first_index, max_iterations, precision) -> Tensor[]

Tensors are borrowed, outputs are allocated on the inputs' device and the work is enqueued on the current stream (the
C-ABI calls underneath take the stream, include/geobipy_b200.h); errors surface as `RuntimeError`
(`_lib.GeobipyB200Error`).  The plain-old-data arguments travel as uint8 CPU tensors holding the bytes of the C structs:
`system` = gbp_fdem_system or gbp_tdem_survey (`pod(struct)`), `options` = gbp_options.  `rjmcmc_run` returns the result
arrays in the order of `RJMCMC_OUTPUTS` (include/geobipy_b200.h gbp_chain_buffers).  Only the CUDA key has an
implementation: there is no CPU path.

The registration is done from Python (`torch.library`) over the ctypes binding rather than with a C++ `TORCH_LIBRARY`
block: the shared library is a plain C-ABI object without any torch symbol in it (the boundary the reference-side binding
of INTEGRATION.md needs), and it stays that way.
"""
import ctypes

import numpy as np
import torch

from . import _lib, ops

__all__ = ["pod", "RJMCMC_OUTPUTS"]

RJMCMC_OUTPUTS = _lib.BUFFER_FIELDS

_LIB = torch.library.Library("geobipy_b200", "DEF")
_LIB.define("fdem_forward(Tensor system, Tensor nlayers, Tensor sigma, Tensor thickness, Tensor altitude, int precision) -> Tensor")
_LIB.define("fdem_sensitivity(Tensor system, Tensor nlayers, Tensor sigma, Tensor thickness, Tensor altitude, int precision) -> (Tensor, Tensor)")
_LIB.define("rjmcmc_run(Tensor system, Tensor options, Tensor data, Tensor altitude, int seed, int first_index, int max_iterations, "
            "int precision) -> Tensor[]")


def pod(struct):
    """The bytes of a C struct (gbp_fdem_system, gbp_tdem_survey, gbp_options) as a uint8 CPU tensor."""
    return torch.frombuffer(bytearray(bytes(struct)), dtype=torch.uint8)


def _system(t):
    raw = t.cpu().numpy().tobytes()
    for cls in (_lib.FdemSystemC, _lib.TdemSurveyC):
        if len(raw) == ctypes.sizeof(cls):
            return cls.from_buffer_copy(raw)
    raise _lib.GeobipyB200Error("system: %d bytes is neither a gbp_fdem_system nor a gbp_tdem_survey" % len(raw))


def _options(t):
    raw = t.cpu().numpy().tobytes()
    if len(raw) != ctypes.sizeof(_lib.OptionsC):
        raise _lib.GeobipyB200Error("options: %d bytes is not a gbp_options" % len(raw))
    return _lib.OptionsC.from_buffer_copy(raw)


def _fdem_forward(system, nlayers, sigma, thickness, altitude, precision):
    return ops.fdem_forward(_system(system), nlayers, sigma.contiguous(), thickness, altitude, precision=precision)


def _fdem_sensitivity(system, nlayers, sigma, thickness, altitude, precision):
    return ops.fdem_forward(_system(system), nlayers, sigma.contiguous(), thickness, altitude, precision=precision, sensitivity=True)


def _rjmcmc_run(system, options, data, altitude, seed, first_index, max_iterations, precision):
    opt = _options(options)
    outs = RJMCMC_OUTPUTS if opt.solve_height else tuple(f for f in RJMCMC_OUTPUTS if f != "height_hist")
    r = ops.rjmcmc_run(_system(system), opt, data, altitude, seed=seed, first_index=first_index, max_iterations=max_iterations,
                       precision=precision, outputs=outs)
    empty = torch.empty((0,), dtype=torch.int32, device=data.device)
    return [r.get(f, empty) for f in RJMCMC_OUTPUTS]


_LIB.impl("fdem_forward", _fdem_forward, "CUDA")
_LIB.impl("fdem_sensitivity", _fdem_sensitivity, "CUDA")
_LIB.impl("rjmcmc_run", _rjmcmc_run, "CUDA")
