"""Helpers to load the reference-generated fixtures (tests/golden/*.npz) and run the oracle on them."""
import os

import numpy as np
import torch

from oracle import gfdn_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
OMNI_CASES = ["omni_n12", "omni_n12_subband_r", "omni_n24"]
SVF_CASES = ["omni_n12_svf", "omni_n12_geq_svf"]  # the second adds GEQ absorption filters (the full-band YAML)
DIR_CASES = ["directional_n27", "directional_n27_skip"]
VARIANT_CASES = ["src_rx_n12", "single_n12", "single_n12_svf"]  # a-8c: source+receiver gains, single position
SRC_RX_SVF_CASES = ["src_rx_n12_svf", "src_rx_n12_svf_in"]  # SVF cascades on both sides / on the source side only


def load(name):
    return dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))


def params_of(g, prefix="param/", requires_grad=False):
    p = {}
    for k, v in g.items():
        if k.startswith(prefix):
            t = torch.tensor(v)
            if t.is_floating_point():
                t = t.to(torch.float64)
                if requires_grad:
                    t.requires_grad_(True)
            p[k[len(prefix):]] = t
    return p


def oracle_omni(g, p, coef_override=None):
    """Forward + trainer loss composition of an omni golden case through the oracle. Returns dict of tensors.
    coef_override: biquad coefficients (B, G, S, 6) to use instead of the ones derived from the MLP (SVF cases)."""
    nfft = int(g["meta/nfft"])
    fs = float(g["meta/fs"])
    z = O.z_grid(nfft, float(g["meta/radius"]))
    delays = torch.tensor(g["meta/delays"], dtype=torch.float64)
    G = p["feedback_loop.M"].shape[0]
    if g["param/delay_filters"].ndim == 4:  # GEQ absorption filters: per-bin complex delay-line gains
        gamma = O.absorption_filter_response(z, torch.tensor(g["param/delay_filters"], dtype=torch.float64))
    else:
        gamma = O.decay_times_to_gain_per_sample(g["meta/t60"], g["meta/delays"], fs, G)
    A = O.coupled_feedback_matrix(p["feedback_loop.M"], p["feedback_loop.alpha"])
    b = p["input_gains"].reshape(-1)
    c = p["output_gains"].reshape(-1)
    d = torch.tensor(g["data/target_early_response"])
    svf = coef = s = None
    if "out/svf_params" in g:  # SVF output filters instead of scalar receiver gains
        svf = O.svf_params_from_mlp(torch.tensor(g["data/listener_position"]), p, int(g["meta/feats"]), G)
        coef = O.svf_to_biquads(svf, O.svf_cutoffs(fs), float(g["meta/pole_factor"]))
        H = O.omni_response_svf(z, delays, gamma, A, b, c, coef if coef_override is None else coef_override, d)
    else:
        s = O.gains_from_mlp(torch.tensor(g["data/norm_listener_position"]), p, int(g["meta/feats"]), G)
        H = O.omni_response(z, delays, gamma, A, b, c, s, d)
    Hs, Hsd = O.sub_fdn_output(z, delays, p["feedback_loop.M"], b, c)
    Huse = H * torch.tensor(g["data/subband_filter"]) if "data/subband_filter" in g else H
    tgt = torch.tensor(g["data/target_rir_response"])
    edc = O.edc_loss(tgt, Huse, float(g["meta/max_ir_len_ms"]), fs)
    edr = O.edr_loss(tgt, Huse)
    spec, spars = O.colorless_losses(Hs, p["feedback_loop.M"], 1.0, 1.0, asym=True)
    total = float(g["meta/edc_w"]) * edc + float(g["meta/edr_w"]) * edr + spec + spars
    return dict(H=H, Huse=Huse, H_sub=Hs, H_sub_per_del=Hsd, A=A, s=s, svf=svf, coef=coef, gamma=gamma, edc=edc, edr=edr, spec=spec,
                spars=spars, total=total, z=z, delays=delays, b=b, c=c, d=d, tgt=tgt)


def oracle_variant(g, p, biquads_from_golden=False):
    """DiffGFDNVarSourceReceiverPos / DiffGFDNSinglePos golden cases (SURVEY a-8c) through the oracle."""
    nfft = int(g["meta/nfft"])
    fs = float(g["meta/fs"])
    z = O.z_grid(nfft, float(g["meta/radius"]))
    delays = torch.tensor(g["meta/delays"], dtype=torch.float64)
    G = p["feedback_loop.M"].shape[0]
    gamma = O.decay_times_to_gain_per_sample(g["meta/t60"], g["meta/delays"], fs, G)
    A = O.coupled_feedback_matrix(p["feedback_loop.M"], p["feedback_loop.alpha"])
    b = p["input_gains"].reshape(-1)
    c = p["output_gains"].reshape(-1)
    d = torch.tensor(g["data/target_early_response"])
    if "data/source_position" in g:
        feats = int(g["meta/feats"])
        def side(tag, svf, pos_gain, pos_svf):
            if not svf:
                return O.gains_from_mlp(torch.tensor(g[pos_gain]), p, feats, G, prefix=f'{tag}_scalars.mlp.model.')
            if biquads_from_golden:  # the reference's own float32 cascades
                coef = torch.tensor(g[f"out/biquads_{'in' if tag == 'input' else 'out'}"])
            else:
                svfp = O.svf_params_from_mlp(torch.tensor(g[pos_svf]), p, feats, G, prefix=f'{tag}_filters.mlp.model.')
                coef = O.svf_to_biquads(svfp, O.svf_cutoffs(fs), float(g["meta/pole_factor"]))
            return O.sos_response(z, coef)
        svf_out = bool(g["meta/svf_out"]) if "meta/svf_out" in g else False
        svf_in = bool(g["meta/svf_in"]) if "meta/svf_in" in g else False
        # receiver side: Gains_from_MLP reads the normalised position, SVF_from_MLP the raw one (gain_filters.py:341, 504)
        s_rx = side("output", svf_out, "data/norm_listener_position", "data/listener_position")
        s_src = side("input", svf_in, "data/source_position", "data/source_position")
        H = O.source_receiver_response(z, delays, gamma, A, b, c, s_rx, s_src, d)
    else:
        def factor(tag, svf):
            if not svf:
                return p[f"{tag}_scalars"].reshape(-1)
            if biquads_from_golden:
                coef = torch.tensor(g[f"out/biquads_{'in' if tag == 'input' else 'out'}"])
            else:
                raw = p[f"{tag}_svf_params"]
                svfp = torch.stack([O.scaled_sigmoid(raw[..., 0], 1e-6, 1.0), O.scaled_sigmoid(raw[..., 1], -6.0, 6.0)], -1)
                coef = O.svf_to_biquads(svfp, O.svf_cutoffs(fs), float(g["meta/pole_factor"]))
            return O.sos_response(z, coef)
        H = O.single_position_response(z, delays, gamma, A, b, c, factor("output", bool(g["meta/svf_out"])),
                                       factor("input", bool(g["meta/svf_in"])), d)
    Hs, Hsd = O.sub_fdn_output(z, delays, p["feedback_loop.M"], b, c)
    tgt = torch.tensor(g["data/target_rir_response"])
    edc = O.edc_loss(tgt, H, float(g["meta/max_ir_len_ms"]), fs)
    edr = O.edr_loss(tgt, H)
    spec, spars = O.colorless_losses(Hs, p["feedback_loop.M"], 1.0, 1.0, asym=True)
    total = float(g["meta/edc_w"]) * edc + float(g["meta/edr_w"]) * edr + spec + spars
    return dict(H=H, H_sub=Hs, A=A, edc=edc, edr=edr, spec=spec, spars=spars, total=total, d=d)


def oracle_directional(g, p):
    nfft = int(g["meta/nfft"])
    fs = float(g["meta/fs"])
    z = O.z_grid(nfft)
    delays = torch.tensor(g["meta/delays"], dtype=torch.float64)
    G = p["feedback_loop.M"].shape[0]
    L = p["feedback_loop.M"].shape[1]
    gamma = O.decay_times_to_gain_per_sample(g["meta/t60"], g["meta/delays"], fs, G)
    A = O.coupled_feedback_matrix(p["feedback_loop.M"], p["feedback_loop.alpha"])
    b = p["input_gains"].reshape(-1)
    c = p["output_gains"].reshape(-1)
    w = O.sh_gains_from_mlp(torch.tensor(g["data/norm_listener_position"]), p, int(g["meta/feats"]), G, L,
                            skip=bool(g["meta/skip"]))
    H_sh = O.directional_response(z, delays, gamma, A, b, c, w)
    Y = torch.tensor(g["data/Y"], dtype=torch.float64)
    H_dir = O.sh_to_directional(H_sh, Y)
    Hs, _ = O.sub_fdn_output(z, delays, p["feedback_loop.M"], b, c)
    env = O.directional_envelopes(np.array([g["meta/t60"]]), float(g["meta/edc_len_ms"]), fs)
    edc = O.directional_edc_loss(H_dir, torch.tensor(g["data/amps"]), env, O.ms_to_samps(float(g["meta/edc_len_ms"]), fs),
                                 O.ms_to_samps(20.0, fs))
    spec, spars = O.colorless_losses(Hs, p["feedback_loop.M"], 1.0, 1.0, asym=False)
    total = float(g["meta/edc_w"]) * edc + spec + spars
    return dict(H_sh=H_sh, H_dir=H_dir, H_sub=Hs, w=w, A=A, env=env, edc=edc, spec=spec, spars=spars, total=total)
