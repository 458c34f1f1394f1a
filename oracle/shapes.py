"""Parameter / buffer shapes of the reference modules on the hot path, derived from the constructor
arguments (TEST INFRASTRUCTURE; mirrors the state_dict contract listed in SURVEY.md section 8a so that
the GPU box, which has no /root/reference, can rebuild recipe weights by name)."""


def _conv(d, p, cout, cin, k, bias):
    d[p + ".weight"] = (cout, cin, k, k)
    if bias:
        d[p + ".bias"] = (cout,)


def _bn(d, p, c):
    for leaf in ("weight", "bias", "running_mean", "running_var"):
        d[f"{p}.{leaf}"] = (c,)
    d[p + ".num_batches_tracked"] = ()


def _res_convblock(d, p, cin, cout, bias=False, norm=True, transpose=False):
    # res_models.py:8-29 (ConvTranspose2d weight is [in, out, k, k])
    d[p + ".conv.weight"] = (cin, cout, 3, 3) if transpose else (cout, cin, 3, 3)
    if bias:
        d[p + ".conv.bias"] = (cout,)
    if norm:
        _bn(d, p + ".norm", cout)


def _resblock(d, p, cin, cout):
    # res_models.py:52-73
    _res_convblock(d, p + ".layers.conv_1", cin, cin)
    _res_convblock(d, p + ".layers.conv_2", cin, cout)
    if cin != cout:
        _conv(d, p + ".projection", cout, cin, 1, True)


def _dual_cell(d, p, c):
    # temporal_ode_bayes.py:77-90 / :225-237
    for i in (1, 2):
        for g in ("update", "reset", "state_tilde"):
            _conv(d, f"{p}.conv_{g}_{i}", c, 2 * c, 3, True)
    _conv(d, p + ".conv_decoder_2", c, c, 3, True)
    t = p + ".trusting_gate.0"
    d[t + ".layers.0.weight"] = (c, 2 * c, 7, 7)
    d[t + ".layers.3.weight"] = (c, c, 1, 1)
    d[t + ".layers.6.weight"] = (c, c, 3, 3)
    for i in (1, 4, 7):
        d[f"{t}.layers.{i}.weight"] = (c,)
        d[f"{t}.layers.{i}.bias"] = (c,)
    d[t + ".projection.0.weight"] = (c, 2 * c, 1, 1)
    d[p + ".trusting_gate.1.weight"] = (2, c, 1, 1)


def nnfo_shapes(c, nf=None):
    """NNFOwithBayesianJumps(input_size=c, hidden_size=c) with FILTER_SIZE nf (temporal_ode_bayes.py:357-393)."""
    nf = nf or c
    d = {}
    p = "p_model.model"
    _resblock(d, p + ".0", c, 2 * c)
    d[p + ".1.fc.0.weight"] = (2 * c // 8, 2 * c)
    d[p + ".1.fc.2.weight"] = (2 * c, 2 * c // 8)
    _resblock(d, p + ".2", 2 * c, 2 * c)
    d[p + ".3.fc.0.weight"] = (2 * c // 8, 2 * c)
    d[p + ".3.fc.2.weight"] = (2 * c, 2 * c // 8)
    _res_convblock(d, p + ".4", 2 * c, 2 * c, bias=True, norm=False)
    _dual_cell(d, "gru_c", c)
    _dual_cell(d, "gru_obs.gru_d", c)
    e = "srvp_encoder"
    for i, (a, b) in enumerate(((c, nf), (nf, 2 * nf), (2 * nf, 2 * nf), (2 * nf, 2 * nf), (2 * nf, 4 * nf))):
        _resblock(d, f"{e}.blocks.{i}", a, b)
    _res_convblock(d, e + ".last_conv.0", 4 * nf, c)
    q = "srvp_decoder"
    _res_convblock(d, q + ".first_upconv", c, 4 * nf, transpose=True)
    for i, (a, b) in enumerate(((4 * nf, 2 * nf), (2 * nf, 2 * nf), (2 * nf, 2 * nf), (2 * nf, nf), (nf, nf))):
        _resblock(d, f"{q}.blocks.{i}", a, b)
    _res_convblock(d, q + ".last_conv.0", nf, nf)
    _res_convblock(d, q + ".last_conv.1", nf, c, bias=True, norm=False, transpose=True)
    return d
