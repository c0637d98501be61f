"""B200 drop-in for ``model/joiner/joiner.py`` of guangkun0818/speech2text.

Same class names, config dataclass, parameter names (``_enc_proj``,
``_pre_proj``, ``_out_projection.{0,1}``: /root/reference/model/joiner/joiner.py:41-55),
call signature and return tuple as the reference (joiner.py:126-182), so
``task_factory/rnnt_task.py:63, 220, 326, 469`` run unchanged.  What changes is
what happens inside ``forward`` on a CUDA device:

  * ``k2.rnnt_loss_smoothed`` / ``k2.get_rnnt_prune_ranges`` (joiner.py:100-117)
    -> libs2t_b200.so (``functional.rnnt_loss_smoothed``, ``get_rnnt_prune_ranges``);
  * ``k2.do_rnnt_pruning`` + add + activation + out-projection (joiner.py:121-123,
    176-178) are NOT executed here in fused mode: ``forward`` returns a
    ``LazyJoinerLogits`` handle in place of the (B,T,R,V) tensor and the loss
    module launches the fused joiner+loss kernels, so the logits never reach HBM.

Environment knobs (no new required config keys -- JoinerConfig(**yaml) must keep working):
  S2T_B200_FUSED=0          materialise logits like the reference (default 1)
  S2T_B200_JOINER_MODE      "fp32" (strict, default) | "bf16" (tensor cores)
  S2T_B200_PRUNE_VARIANT    "B" (default: upstream's get_rnnt_prune_ranges since early 2023, what the reference's
                            k2 pins resolve to) | "A" (the older sliding-window function)
"""
from __future__ import annotations

import dataclasses
import os
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from . import functional as F2


@dataclasses.dataclass
class JoinerConfig:
    """ Joiner Config interface (field-for-field the reference's, joiner.py:16-26) """
    input_dim: int  # Input dimension of encoder_out and predictor_out.
    output_dim: int  # Output dimension, refered as vocab size
    inner_dim: int = 256  # Inner dim of last projection layer
    activation: str = "relu"  # activation func, choose from ("relu", "tanh")
    prune_range: int = 5  # specify as -1 if pruned rnnt loss not applied
    lm_scale: float = 0.0  # lm_scale applied in simple_loss of pruned_rnnt
    am_scale: float = 0.0  # am_scale applied in simple_loss of pruned_rnnt
    use_out_project: bool = True  # If apply last output projection, if false, params saved


def _mode_from_env() -> int:
    mode = os.environ.get("S2T_B200_JOINER_MODE", "fp32").lower()
    if mode in ("fp32", "simt"):
        return _lib.MODE_FP32_SIMT
    if mode in ("bf16", "tc"):
        return _lib.MODE_BF16_TC
    raise ValueError(f"S2T_B200_JOINER_MODE must be fp32 or bf16, got {mode}")


class LazyJoinerLogits:
    """Stand-in for the joiner's (B, T, R, V) output tensor.

    Holds what is needed to evaluate ``W2 (W1 act(am + lm[ranges]) + b1) + b2``
    tile by tile inside the loss kernels.  It quacks like the tensor for the few
    things the reference's tasks and tests do with ``joiner_out`` (``.shape``,
    ``.to(dtype)``, ``.float()``); anything else should call ``materialize()``.
    """

    def __init__(self, am, lm, W1, b1, W2, b2, ranges, act: int, mode: int):
        self.am, self.lm = am, lm
        self.W1, self.b1, self.W2, self.b2 = W1, b1, W2, b2
        self.ranges = ranges
        self.act = act
        self.mode = mode

    @property
    def shape(self) -> torch.Size:
        B, T, V = self.am.shape
        R = self.ranges.shape[2] if self.ranges is not None else self.lm.shape[1]
        return torch.Size((B, T, R, V))

    def size(self, dim: Optional[int] = None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self) -> int:
        return 4

    @property
    def dtype(self):
        return torch.float32

    @property
    def device(self):
        return self.am.device

    def to(self, *args, **kwargs):
        return self

    def float(self):
        return self

    def materialize(self) -> torch.Tensor:
        """fp32 logits (no autograd); costs B*T*R*V*4 bytes."""
        return F2.joiner_materialize(self.am, self.lm, self.W1, self.b1, self.W2, self.b2, self.ranges,
                                     self.act, _lib.MODE_FP32_SIMT)


class Joiner(nn.Module):
    """ Joiner of both Predictor and Encoder of Rnnt """

    def __init__(self, config: JoinerConfig) -> None:
        super(Joiner, self).__init__()

        self._input_dim = config.input_dim
        self._output_dim = config.output_dim
        self._inner_dim = config.inner_dim

        self._enc_proj = nn.Linear(self._input_dim, self._output_dim, bias=True)
        self._pre_proj = nn.Linear(self._input_dim, self._output_dim, bias=True)

        if config.activation == "relu":
            self._activation = nn.ReLU()
            self._act_code = _lib.ACT_RELU
        elif config.activation == "tanh":
            self._activation = nn.Tanh()
            self._act_code = _lib.ACT_TANH
        else:
            raise ValueError(f"Unsupported activation {config.activation}")

        self._use_out_project = config.use_out_project
        if self._use_out_project:
            self._out_projection = nn.Sequential(nn.Linear(self._output_dim, self._inner_dim),
                                                 nn.Linear(self._inner_dim, self._output_dim))
        else:
            self._out_projection = nn.Identity()  # Placeholder

        self._log_softmax = nn.LogSoftmax(dim=-1)

        self._blank_token = 0  # 0 is strictly set for both Ctc and Rnnt.
        self._prune_range = config.prune_range
        self._lm_scale = config.lm_scale
        self._am_scale = config.am_scale

    @property
    def prune_range(self) -> int:
        return self._prune_range

    @property
    def blank_token(self) -> int:
        return self._blank_token

    @torch.jit.unused
    def _out_proj_params(self):
        if not self._use_out_project:
            return None, None, None, None
        l0, l1 = self._out_projection[0], self._out_projection[1]
        return l0.weight, l0.bias, l1.weight, l1.bias

    @torch.jit.unused
    def _simple_loss_and_ranges(self, am: torch.Tensor, encoder_out_lengths: torch.Tensor, lm: torch.Tensor,
                                target_lengths: torch.Tensor, target: torch.Tensor, mode: int = 0, row_max=None,
                                prepared_ws=None):
        """joiner.py:74-117 without the pruning gather: boundary, simple loss, ranges."""
        boundary = F2.make_boundary(target_lengths, encoder_out_lengths, am.device)
        assert len(target.shape) == 2  # (B, U)
        assert lm.shape[-1] >= self._output_dim and am.shape[-1] >= self._output_dim, (
            "If pruned rnnt loss applied, output dim of encoder and predictor should be mandatorily "
            "larger than num of tokens. Please check your config")
        # Pruned rnnt loss strictly required fp32
        simple_loss, (px_grad, py_grad) = F2.rnnt_loss_smoothed(
            lm=lm.to(dtype=torch.float32),
            am=am.to(dtype=torch.float32),
            symbols=target.to(am.device),
            termination_symbol=self.blank_token,
            lm_only_scale=self._lm_scale,
            am_only_scale=self._am_scale,
            boundary=boundary,
            reduction="mean",
            return_grad=True,
            mode=mode,
            row_max=row_max,
            prepared_ws=prepared_ws,
        )
        ranges = F2.get_rnnt_prune_ranges(px_grad=px_grad, py_grad=py_grad, boundary=boundary,
                                          s_range=self.prune_range)
        return boundary, ranges, simple_loss

    @torch.jit.unused
    def forward(
        self,
        encoder_out: torch.Tensor,
        encoder_out_lengths: torch.Tensor,
        predict_out: torch.Tensor,
        target_lengths: torch.Tensor,
        target: torch.Tensor = torch.empty(0, 0)
    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """ Args and returns exactly as the reference (joiner.py:126-182):
                encoder_out: (B, T, D); encoder_out_lengths: (B)
                predict_out: (B, U + 1, D); target_lengths: (B)
                target: (B, U), used when the pruned loss is applied.
            Returns (output, boundary, ranges, simple_loss); in fused mode
            ``output`` is a LazyJoinerLogits handle of shape (B, T, R, V).
        """
        if not encoder_out.is_cuda:
            raise _lib.S2TError("speech2text_b200.Joiner.forward needs CUDA tensors: this build has no "
                                "CPU path (streaming_step and the ONNX exports are plain torch).")
        # Project both encoder_out and predictor_out into vocab_size
        mode = _mode_from_env()
        prepared_ws = None
        if mode == _lib.MODE_BF16_TC and os.environ.get("S2T_B200_PROJ_TC", "1") != "0":
            # two aliases of each projection: one for the simple loss, one for the joiner (functional._LinearTC)
            # ... and, when the simple loss follows, the row maxima it needs as a by-product of the GEMM epilogue
            want = self.prune_range > 0
            # the predictor-side projection (under one wave of tiles) runs on a side stream next to the encoder-side one,
            # and so does the lm side of the simple loss's normaliser, which needs nothing else
            side = os.environ.get("S2T_B200_PROJ_OVERLAP", "1") != "0" and not _lib.profiling()
            lm, lm_j, *lm_max = F2.linear_tc_pair(predict_out, self._pre_proj.weight, self._pre_proj.bias, row_max=want,
                                                  on_side_stream=side)
            if want and side and target.dim() == 2:
                prepared_ws = F2.simple_loss_prepare_lm(lm, lm_max[0], target, encoder_out.shape[1], self.blank_token, mode,
                                                        on_side_stream=True)
            am, am_j, *am_max = F2.linear_tc_pair(encoder_out, self._enc_proj.weight, self._enc_proj.bias, row_max=want)
            if side:
                F2.join_side_stream(encoder_out.device)
            row_max = (am_max[0], lm_max[0]) if want else None
        else:
            am = am_j = self._enc_proj(encoder_out)
            lm = lm_j = self._pre_proj(predict_out)
            row_max = None

        if self.prune_range > 0:
            assert target.shape[0] == target_lengths.shape[0]
            boundary, ranges, simple_loss = self._simple_loss_and_ranges(am, encoder_out_lengths, lm,
                                                                         target_lengths, target, mode, row_max, prepared_ws)
        else:
            # For API consistency
            boundary = None
            ranges = None
            simple_loss = None

        W1, b1, W2, b2 = self._out_proj_params()
        fused = os.environ.get("S2T_B200_FUSED", "1") != "0"
        if fused:
            output = LazyJoinerLogits(am_j, lm_j, W1, b1, W2, b2, ranges, self._act_code, mode)
        else:
            # the reference's own materialising ops (joiner.py:121-123, 166-178)
            if ranges is not None:
                B, T, R = ranges.shape
                S1, C = lm.shape[1], lm.shape[2]
                am_p = am_j.unsqueeze(2).expand((B, T, R, am.shape[-1]))
                lm_p = torch.gather(lm_j.unsqueeze(1).expand((B, T, S1, C)), dim=2,
                                    index=ranges.reshape((B, T, R, 1)).expand((B, T, R, C)))
            else:
                am_p = am_j.unsqueeze(2).contiguous()
                lm_p = lm_j.unsqueeze(1).contiguous()
            output = self._out_projection(self._activation(am_p + lm_p))

        # Use raw output of joiner for rnnt_loss compute since log_softmax will be
        # done within rnnt_loss.
        return output, boundary, ranges, simple_loss

    @torch.jit.export
    @torch.inference_mode(mode=True)
    def streaming_step(self, encoder_out: torch.Tensor, predictor_out: torch.Tensor):
        # Streaming inference step (joiner.py:184-207): 1 encoder frame, beam_size predictor
        # states; plain torch ops so that torch.jit.script(joiner) keeps working.
        assert encoder_out.shape[0] == 1 and encoder_out.shape[1] == 1
        assert predictor_out.shape[1] == 1

        encoder_out = self._enc_proj(encoder_out)
        predictor_out = self._pre_proj(predictor_out)

        encoder_out = encoder_out.unsqueeze(2).contiguous()
        predictor_out = predictor_out.unsqueeze(1).contiguous()

        joint_encodings = encoder_out + predictor_out
        activation_out = self._activation(joint_encodings)
        output = self._out_projection(activation_out)

        output = self._log_softmax(output)  # (beam, 1, 1, V)
        output = output.squeeze(1).squeeze(1)  # (beam, V)
        return output

    @torch.jit.export
    @torch.inference_mode(mode=True)
    def sherpa_onnx_streaming_step(self, encoder_out: torch.Tensor, predictor_out: torch.Tensor):
        # Wrapped forward for Onnx export (joiner.py:209-221)
        encoder_out = self._enc_proj(encoder_out)  # (N, V)
        predictor_out = self._pre_proj(predictor_out)  # (N, V)
        joint_encodings = encoder_out + predictor_out
        activation_out = self._activation(joint_encodings)
        output = self._out_projection(activation_out)
        return output

    def onnx_export(self, export_path, for_mnn=True, for_sherpa=True):
        """ Interface for onnx export (joiner.py:223-228). """
        if for_sherpa:
            self._sherpa_onnx_export(export_path=export_path)
        if for_mnn:
            self._mnn_onnx_export(export_path=export_path)

    def _mnn_onnx_export(self, export_path):
        """ Export Onnx model for mnn deploy: fixed shapes, beam 11 (joiner.py:230-254). """
        export_filename = os.path.join(export_path, "joiner.onnx")
        self.train(False)
        restore = self.forward
        self.forward = self.streaming_step
        try:
            enc_out = torch.rand(1, 1, self._input_dim, dtype=torch.float32)
            pred_out = torch.rand(11, 1, self._input_dim, dtype=torch.float32)
            torch.onnx.export(self, (enc_out, pred_out), export_filename, verbose=True, opset_version=13,
                              input_names=["enc_out", "pred_out"], output_names=["logit"])
        finally:
            self.forward = restore

    def _sherpa_onnx_export(self, export_path):
        """ Export Onnx model for sherpa-onnx: dynamic batch axis + joiner_dim metadata
            (joiner.py:256-296). """
        export_filename = os.path.join(export_path, "joiner.onnx")
        self.train(False)
        restore = self.forward
        self.forward = self.sherpa_onnx_streaming_step
        try:
            ts_joiner = torch.jit.script(self)
            enc_out = torch.rand(11, self._input_dim, dtype=torch.float32)
            pre_out = torch.rand(11, self._input_dim, dtype=torch.float32)
            torch.onnx.export(ts_joiner, (enc_out, pre_out), export_filename, verbose=False, opset_version=13,
                              input_names=["encoder_out", "decoder_out"], output_names=["logit"],
                              dynamic_axes={"encoder_out": {0: "N"}, "decoder_out": {0: "N"}, "logit": {0: "N"}})
            self._add_meta_data(filename=export_filename, meta_data={"joiner_dim": str(self._input_dim)})
        finally:
            self.forward = restore

    def _add_meta_data(self, filename, meta_data):
        """ Add meta data to an ONNX model in place (joiner.py:298-310). """
        import onnx  # optional dependency, only needed for the export helpers
        model = onnx.load(filename)
        for key, value in meta_data.items():
            meta = model.metadata_props.add()
            meta.key = key
            meta.value = value
        onnx.save(model, filename)
