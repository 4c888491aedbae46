"""Lazy feature handle: keeps the reference's own call sequence on ONE fused kernel.

The reference's callers (utils/tracker.py:297-335, utils/mesher.py:112-130) do

    geo, col, w, cnt, cert = neural_points.query_feature(coord, ...)     # kNN + gathers + IDW
    sdf  = sdf_mlp.sdf(geo)                                              # decoder
    grad = get_gradient(coord, sdf)                                      # autograd (utils/tools.py:247-260)

With `weighted_first` the feature tensor is only ever fed to a decoder, so `query_feature` of the drop-in returns a
`FusedFeatures` handle instead of materialising [N, F+3] floats; `Decoder.sdf` / `Decoder.regress_color` recognise
it and run the fused K1 kernel (search + gather + decoder + analytic d/dq) through a `torch.autograd.Function`
whose backward hands the kernel's gradient to `torch.autograd.grad`.  Anything else that touches the handle
(`Decoder.mlp`, torch functions, attribute access) materialises the tensor with the differentiable torch ops of
`NeuralPoints._query_feature_eager`, so unknown callers keep working, just not fused."""
import torch


class _Group:
    """State shared by the geo and colour handles of one query_feature call."""

    def __init__(self, npm, query_points, query_locally):
        self.npm, self.q, self.query_locally = npm, query_points, query_locally
        self.sdf_decoder = None
        self.eager = None  # (geo, col) materialised on demand

    def materialise(self):
        if self.eager is None:
            geo, col, _, _, _ = self.npm._query_feature_eager(self.q, None, False, self.query_locally, True,
                                                              self.npm.color_features is not None)
            self.eager = (geo, col)
        return self.eager


class _FusedDecode(torch.autograd.Function):
    """value [N] (or [N, C]) with d value / d query taken from the kernel."""

    @staticmethod
    def forward(ctx, q, group, decoder, which):
        npm = group.npm
        need = bool(ctx.needs_input_grad[0])
        if which == "geo":
            o = npm.query_sdf(q.detach().contiguous(), decoder, query_locally=group.query_locally, need_grad=need)
            val = o["sdf"].clone()
            jac = o["grad"].clone() if need else None  # [N, 3]
        else:
            o = npm.query_sdf(q.detach().contiguous(), group.sdf_decoder, query_locally=group.query_locally,
                              need_grad=need, color_decoder=decoder, color_grad=need)
            val = o["color"].clone()
            jac = o["color_grad"].clone() if need else None  # [N, C, 3]
        ctx.which = which
        ctx.save_for_backward(jac) if need else None
        ctx.has_jac = need
        return val

    @staticmethod
    def backward(ctx, gout):
        if not ctx.has_jac:
            return None, None, None, None
        (jac,) = ctx.saved_tensors
        if ctx.which == "geo":
            return gout.unsqueeze(1) * jac, None, None, None
        return (gout.unsqueeze(2) * jac).sum(1), None, None, None


class FusedFeatures:
    """Stands in for the [N, F+3] feature tensor of a weighted_first query."""

    def __init__(self, group, which):
        self._group, self._which = group, which

    def materialize(self):
        geo, col = self._group.materialise()
        return geo if self._which == "geo" else col

    # ---- the fused consumers ------------------------------------------------------------------
    def decode(self, decoder, color: bool):
        """Decoder.sdf / Decoder.regress_color on the handle; None when the fused kernel does not apply."""
        g = self._group
        if any(p.requires_grad for p in decoder.parameters()) and torch.is_grad_enabled():
            return None  # training the decoder needs the torch graph through its weights
        if not color:
            if self._which != "geo":
                return None
            g.sdf_decoder = decoder
            return _FusedDecode.apply(g.q, g, decoder, "geo")
        if self._which != "color" or g.sdf_decoder is None:
            return None
        return _FusedDecode.apply(g.q, g, decoder, "color")

    # ---- everything else: behave like the materialised tensor -----------------------------------
    def __getattr__(self, name):
        return getattr(self.materialize(), name)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        conv = lambda a: a.materialize() if isinstance(a, FusedFeatures) else a  # noqa: E731
        args = tuple(conv(a) for a in args)
        kwargs = {k: conv(v) for k, v in (kwargs or {}).items()}
        return func(*args, **kwargs)
