"""Drop-in `Decoder` (reference: model/decoder.py:14 of PRBonn/PIN_SLAM).

Same constructor, submodule names (`layers`, `lout` -> identical state_dict keys,
so reference checkpoints load), `sdf_scale`, and methods.  The methods themselves
stay plain torch modules -- they are what unchanged reference callers invoke on
materialised feature tensors.  The fused CUDA path never calls them: it reads the
weights through `handle()` (a pinb200_decoder_view over the parameter storage) and
evaluates the MLP inside the K1 / K2 kernels."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .fused_features import FusedFeatures


class Decoder(nn.Module):
    def __init__(self, config, hidden_dim, hidden_level, out_dim, is_time_conditioned=False):
        super().__init__()
        if is_time_conditioned:
            raise NotImplementedError("time-conditioned decoders are unused by PIN-SLAM and not supported")
        if getattr(config, "pos_encoding_band", 0) > 0:
            raise NotImplementedError("pin_slam_b200: positional encoding is not supported")
        self.out_dim = out_dim
        self.use_leaky_relu = config.mlp_leaky_relu
        bias_on = config.mlp_bias_on
        in_dim = config.feature_dim + config.pos_input_dim
        widths = [in_dim] + [hidden_dim] * hidden_level
        self.layers = nn.ModuleList(nn.Linear(a, b, bias_on) for a, b in zip(widths[:-1], widths[1:]))
        self.lout = nn.Linear(hidden_dim, out_dim, bias_on)
        self.sdf_scale = 1.0
        if config.main_loss_type == "bce":
            self.sdf_scale = config.logistic_gaussian_ratio * config.sigma_sigmoid_m
        self.to(config.device)
        self._handle = {}

    # ---- torch path for materialised features (API compatibility) ----
    def mlp(self, features):
        if isinstance(features, FusedFeatures):
            features = features.materialize()
        act = F.leaky_relu if self.use_leaky_relu else F.relu
        h = features
        for layer in self.layers:
            h = act(layer(h))
        return self.lout(h)

    def sdf(self, features):
        if isinstance(features, FusedFeatures):  # lazy handle of a weighted_first query: one fused K1 launch
            fused = features.decode(self, color=False)
            if fused is not None:
                return fused
        return self.mlp(features).squeeze(1) * self.sdf_scale

    def time_conditionded_sdf(self, features, ts):
        k = features.shape[1]
        return self.sdf(torch.cat((features, ts.repeat(k).view(-1, k, 1)), dim=-1))

    def occupancy(self, features):
        return torch.sigmoid(self.sdf(features) / -self.sdf_scale)

    def sem_label_prob(self, features):
        return F.log_softmax(self.mlp(features), dim=-1)

    def sem_label(self, features):
        return torch.argmax(self.sem_label_prob(features), dim=1)

    def regress_color(self, features):
        if isinstance(features, FusedFeatures):
            fused = features.decode(self, color=True)
            if fused is not None:
                return fused
        return torch.sigmoid(self.mlp(features))

    # ---- kernel view ----
    def handle(self, sigmoid_out: bool = False) -> ops.DecoderHandle:
        """pinb200_decoder_view over the live parameter storage (rebuilt if a parameter was re-allocated)."""
        ps = self._handle.get("ps")
        if ps is None or ps[-2] is not self.lout.weight:  # Parameter objects are stable; their storage may move
            ps = [p for l in self.layers for p in (l.weight, l.bias)] + [self.lout.weight, self.lout.bias]
        key = (sigmoid_out,) + tuple(0 if p is None else p.data_ptr() for p in ps)
        h = self._handle.get("h")
        if h is None or self._handle.get("key") != key:
            h = ops.DecoderHandle([l.weight.data for l in self.layers],
                                  [None if l.bias is None else l.bias.data for l in self.layers], self.lout.weight.data,
                                  None if self.lout.bias is None else self.lout.bias.data,
                                  out_scale=1.0 if sigmoid_out else self.sdf_scale, leaky=self.use_leaky_relu,
                                  sigmoid_out=sigmoid_out)
            self._handle = {"h": h, "key": key, "ps": ps}
        return h

    def flat_parameters(self) -> torch.Tensor:
        """All parameters as ONE contiguous vector [w0|b0|w1|b1|...|w_out|b_out] (the layout K2/K3 use).
        On first use the nn.Linear parameters are re-pointed to views of this vector, so the fused Adam
        step updates the modules in place and state_dict()/checkpoints keep working."""
        ps = [p for l in self.layers for p in (l.weight, l.bias) if p is not None]
        ps += [p for p in (self.lout.weight, self.lout.bias) if p is not None]
        flat = getattr(self, "_flat", None)
        ok = flat is not None and flat.device == ps[0].device
        if ok:
            off = 0
            for p in ps:
                ok = ok and p.data_ptr() == flat.data_ptr() + 4 * off
                off += p.numel()
        if not ok:
            flat = torch.cat([p.detach().reshape(-1) for p in ps]).contiguous()
            off = 0
            for p in ps:
                p.data = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            self._flat = flat
            self._handle = {}
        return flat

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_handle"] = {}
        state.pop("_flat", None)
        return state
