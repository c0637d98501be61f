"""B200 drop-in for ``model/predictor/stateless_predictor.py`` of guangkun0818/speech2text (SURVEY.md 8 row f-3:
the producer of ``predict_out``, the step right before the joiner).

Same config dataclass, parameter names (``_embedding``, ``_conv``, ``_output_linear``:
/root/reference/model/predictor/stateless_predictor.py:37-53), ``forward`` / ``init_state`` /
``streaming_step`` signatures and return tuples, so ``model/predictor/predictor.py`` (the factory) and
``task_factory/rnnt_task.py:464-466`` run unchanged.  On a CUDA device ``forward`` replaces the reference's
embedding -> transpose -> depthwise Conv1d -> transpose chain (stateless_predictor.py:90-97) with ONE fused
gather-multiply-add kernel (``s2t_predictor_embed_conv_fwd/bwd``) and runs the output Linear on the tensor-core
GEMM in tensor-core mode.  ``streaming_step`` (one token, used by decoding) and the ONNX exports are plain torch.
"""
from __future__ import annotations

import dataclasses
import os
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from . import functional as F2


@dataclasses.dataclass
class StatelessPredictorConfig:
    """ Stateless Predictor Config (field-for-field the reference's, stateless_predictor.py:19-25) """
    num_symbols: int = 128  # <blank_id> = 0, <sos> = num_symbols - 1 are taken into account
    output_dim: int = 1024
    symbol_embedding_dim: int = 512
    context_size: int = 5


class StatelessPredictor(nn.Module):
    """ Stateless Predictor: Embedding + depthwise Conv1d over the last ``context_size`` tokens + Linear """

    def __init__(self, config: StatelessPredictorConfig) -> None:
        super(StatelessPredictor, self).__init__()
        self._sos_token = config.num_symbols - 1
        self._blank_token = 0  # 0 is strictly set for both Ctc and Rnnt.
        self._embedding_dim = config.symbol_embedding_dim
        self._num_symbols = config.num_symbols
        self._embedding = nn.Embedding(num_embeddings=self._num_symbols, embedding_dim=self._embedding_dim)
        assert config.context_size >= 1, "context_size should be greater than or eq to 1"
        self._context_size = config.context_size
        self._output_dim = config.output_dim
        self._conv = nn.Conv1d(in_channels=self._embedding_dim, out_channels=self._embedding_dim,
                               kernel_size=self._context_size, stride=1, padding=0, groups=self._embedding_dim,
                               bias=False)
        self._output_linear = nn.Linear(self._embedding_dim, self._output_dim)

    @property
    def sos_token(self) -> int:
        return self._sos_token

    @property
    def blank_token(self) -> int:
        return self._blank_token

    def _left_padding(self, x: torch.Tensor) -> torch.Tensor:
        # tokens left-padded with <blank_id> (stateless_predictor.py:63-71)
        assert len(x.shape) == 2  # (B, U)
        return F.pad(x.float(), (1, 0, 0, 0), value=float(self.blank_token)).to(torch.int32)  # (B, 1 + U)

    @torch.jit.unused
    def forward(self, input: torch.Tensor, lengths: torch.Tensor,
                state: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """ Training graph (stateless_predictor.py:74-99).
            input: tokens (B, U); lengths (B); state: init state (1, context_size - 1).
            Returns (output (B, 1 + U, output_dim), lengths, out_state (B, context_size)). """
        if not input.is_cuda:
            raise _lib.S2TError("speech2text_b200.StatelessPredictor.forward needs CUDA tensors: this build has no "
                                "CPU path (streaming_step and the ONNX exports are plain torch).")
        bs = input.shape[0]
        state = state.repeat(bs, 1).to(input.device)  # (B, context_size - 1)
        ctxed_input = torch.concat([state, self._left_padding(input)], dim=1)  # (B, context_size + U)
        cache_pos = ctxed_input.shape[1] - self._context_size
        out_state = ctxed_input[:, cache_pos:]
        h = F2.predictor_embed_conv(ctxed_input, self._embedding.weight, self._conv.weight)  # (B, 1 + U, E)
        if os.environ.get("S2T_B200_JOINER_MODE", "fp32").lower() in ("bf16", "tc"):
            output = F2.linear_tc(h, self._output_linear.weight, self._output_linear.bias)
        else:
            output = self._output_linear(h)
        return output, lengths, out_state

    @torch.jit.export
    def init_state(self, batch_size: int = 1) -> torch.Tensor:
        # [blank, ..., blank] of length context_size - 1
        return torch.zeros(batch_size, self._context_size - 1).to(torch.int32)

    def _context_layer(self, tokens: torch.Tensor) -> torch.Tensor:
        embs = self._embedding(tokens).contiguous().transpose(1, 2)  # (B, U, E) -> (B, E, U)
        return self._output_linear(self._conv(embs).contiguous().transpose(1, 2))

    @torch.jit.export
    @torch.inference_mode(mode=True)
    def streaming_step(self, input: torch.Tensor, state: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        # one token per call (stateless_predictor.py:107-124)
        assert input.shape[1] == 1  # only support sequence length = 1
        ctxed_input = torch.concat([state.to(input.device), input], dim=1)
        cache_pos = ctxed_input.shape[1] - self._context_size + 1
        out_state = ctxed_input[:, cache_pos:]
        return self._context_layer(ctxed_input), out_state

    @torch.jit.export
    @torch.inference_mode(mode=True)
    def sherpa_onnx_streaming_step(self, input: torch.Tensor):
        # wrapped forward for ONNX export (stateless_predictor.py:126-136)
        assert input.shape[1] == self._context_size
        return self._context_layer(input).squeeze(1)  # (B, D)

    def onnx_export(self, export_path, for_mnn=True, for_sherpa=True):
        """ Interface for onnx export (stateless_predictor.py:138-143). """
        if for_sherpa:
            self._sherpa_onnx_export(export_path=export_path)
        if for_mnn:
            self._mnn_onnx_export(export_path=export_path)

    def _export_with(self, fn, args, filename, **kw):
        self.train(False)
        restore = self.forward
        self.forward = fn
        try:
            torch.onnx.export(self, args, filename, opset_version=13, **kw)
        finally:
            self.forward = restore

    def _mnn_onnx_export(self, export_path):
        """ init-state model + streaming-step model with fixed shapes (stateless_predictor.py:145-190). """
        self._export_with(self.init_state, 10, os.path.join(export_path, "predictor_init.onnx"), verbose=True,
                          input_names=["beam_size"], output_names=["states"])
        batch_size = 10
        self._export_with(self.streaming_step, (torch.randint(1, 128, (batch_size, 1)), self.init_state(batch_size)),
                          os.path.join(export_path, "predictor.onnx"), verbose=True,
                          input_names=["pred_in", "prev_states"], output_names=["pred_out", "next_states"])

    def _sherpa_onnx_export(self, export_path):
        """ dynamic batch axis + context_size / vocab_size metadata (stateless_predictor.py:192-229). """
        export_filename = os.path.join(export_path, "predictor.onnx")
        self._export_with(self.sherpa_onnx_streaming_step, torch.zeros(10, self._context_size, dtype=torch.int64),
                          export_filename, verbose=False, input_names=["y"], output_names=["decoder_out"],
                          dynamic_axes={"y": {0: "N"}, "decoder_out": {0: "N"}})
        self._add_meta_data(filename=export_filename,
                            meta_data={"context_size": str(self._context_size), "vocab_size": str(self._num_symbols)})

    def _add_meta_data(self, filename, meta_data):
        import onnx  # optional dependency, only needed for the export helpers
        model = onnx.load(filename)
        for key, value in meta_data.items():
            meta = model.metadata_props.add()
            meta.key = key
            meta.value = value
        onnx.save(model, filename)
