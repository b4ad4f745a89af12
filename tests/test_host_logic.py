"""CPU tests of the host side: C-ABI surface, parameter tree / checkpoint compatibility, the flat
arena, optimizer grouping, adapter bookkeeping. No kernel runs here (there is no GPU in the build
container); compute parity lives in the -m gpu tests."""
import copy
import ctypes
import os
import re

import pytest
import torch

from oracle import vilt_oracle as vo
from tests.golden_util import ALL_TASKS, TINY

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _learner(dims=TINY, tasks=ALL_TASKS):
    from climb_b200.modeling import B200ViltConfig, B200ViltContinualLearner, B200ViltEncoderWrapper, B200ViltModel
    cfg = B200ViltConfig(hidden_size=dims.hidden_size, num_hidden_layers=dims.num_hidden_layers,
                         num_attention_heads=dims.num_attention_heads, intermediate_size=dims.intermediate_size,
                         image_size=dims.image_size, patch_size=dims.patch_size, vocab_size=dims.vocab_size,
                         max_position_embeddings=dims.max_position_embeddings)
    enc = B200ViltEncoderWrapper(None, B200ViltModel(cfg), torch.device("cpu"))
    return B200ViltContinualLearner(list(tasks), enc, dims.hidden_size, vo.TASK_SPECS)


def test_library_exports_every_declared_symbol():
    from climb_b200 import _lib
    header = open(os.path.join(ROOT, "include", "climb_b200.h")).read()
    declared = set(re.findall(r"\b(climb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 19
    for name in sorted(declared):
        assert hasattr(_lib.lib, name), f"{name} declared in include/climb_b200.h but not exported"
    assert _lib.climb_version() >= 100


def test_struct_layouts_match_header_sizes():
    """ctypes mirrors of the C structs: field counts / sizes that would silently corrupt calls."""
    from climb_b200 import _lib
    assert ctypes.sizeof(_lib.ViltLayerC) == 20 * 8 + 8
    assert ctypes.sizeof(_lib.AdamWChunkC) == 16
    assert ctypes.sizeof(_lib.ViltDimsC) == 14 * 4          # ..., vocab_size, type_vocab_size, precision, hidden_dropout, attn_dropout
    assert ctypes.sizeof(_lib.GemmDesc) % 8 == 0
    assert ctypes.sizeof(_lib.ViltBatchC) == 4 * 4 + 6 * 8 + 8 + 8 + 8 + 8 + 8 + 8  # ..., image_type_idx_scalar(+pad), patch_geom, n_patch_slots + training, dropout_seed, image_repeat(+pad), patch_select
    assert ctypes.sizeof(_lib.ViltParamsC) == 14 * 8 + 8 + 4 * 4 + 8              # offsets, layer*, adapter_r/act + flags, shadow_lo
    assert ctypes.sizeof(_lib.BertDimsC) == 5 * 4
    assert ctypes.sizeof(_lib.BertLayerC) == 12 * 8
    assert ctypes.sizeof(_lib.BertParamsC) == 5 * 8 + 8
    assert ctypes.sizeof(_lib.BertBatchC) == 2 * 4 + 3 * 8


def test_cpu_tensors_are_rejected_loudly():
    from climb_b200 import _lib
    with pytest.raises(_lib.ClimbError):
        _lib.ptr(torch.zeros(4))
    learner = _learner()
    batch = vo.synth_batch("vqa", 2, TINY, T=8, image_hw=(32, 32))
    with pytest.raises(_lib.ClimbError):
        learner.forward_tensors("vqa", {k: v for k, v in batch.items() if k != "target"})


def test_state_dict_is_key_and_shape_compatible_with_the_reference():
    learner = _learner()
    shapes = vo.param_shapes(TINY, ALL_TASKS)
    sd = learner.state_dict()
    ours = {k: tuple(v.shape) for k, v in sd.items() if "position_ids" not in k}
    assert ours == dict(shapes)
    assert "vilt_encoder.vilt.embeddings.text_embeddings.position_ids" in sd
    # registration order = the reference's named_parameters() order (EWC / optimizers iterate it)
    assert [n for n, _ in learner.named_parameters()] == list(shapes.keys())
    base = vo.param_shapes(vo.ViltDims(), ALL_TASKS)
    assert sum(int(torch.tensor(s).prod()) for s in base.values()) == 121_145_919      # SURVEY.md 8c (includes the third modality row)


def test_adapter_parameter_names_and_freezing():
    learner = _learner()
    learner.add_adapter("nlvr2", "houlsby")
    learner.add_adapter("vqa", {"reduction_factor": 4, "non_linearity": "relu", "mh_adapter": False, "output_adapter": True})
    names = [n for n, _ in learner.named_parameters()]
    exp = vo.param_shapes(TINY, ALL_TASKS, adapters={"nlvr2": 8})
    assert set(n for n in exp if ".adapters.nlvr2." in n) <= set(names)
    assert not any(".attention.output.adapters.vqa." in n for n in names)       # pfeiffer-style: output only
    learner.train_adapter("nlvr2")
    assert learner.get_active_adapters() == "nlvr2"
    for n, p in learner.named_parameters():
        expect = (".adapters.nlvr2." in n) or n.startswith("task_layer.")
        assert p.requires_grad == expect, n
    with pytest.raises(NotImplementedError):
        learner.add_adapter("snli-ve", {"reduction_factor": 16, "phm_layer": True})
    with pytest.raises(ValueError):
        learner.set_active_adapters("missing")


def test_arena_views_qkv_contiguity_and_rebinding():
    learner = _learner()
    vilt = learner.vilt_encoder.vilt
    sd = vo.synth_state_dict(TINY, ALL_TASKS, seed=5)
    learner.load_state_dict(sd, strict=False)
    arena = vilt._arena
    assert arena.sync(allow_cpu=True) is True
    assert arena.sync(allow_cpu=True) is False
    off = arena.offsets
    d = TINY.hidden_size
    q = "encoder.layer.1.attention.attention."
    assert off[q + "key.weight"] == off[q + "query.weight"] + d * d
    assert off[q + "value.weight"] == off[q + "key.weight"] + d * d
    assert off[q + "key.bias"] == off[q + "query.bias"] + d
    assert all(o % 64 == 0 for o in off.values())
    # parameters are views: the fused [3d, d] weight is readable straight from theta
    fused = arena.theta[off[q + "query.weight"]: off[q + "query.weight"] + 3 * d * d].view(3 * d, d)
    assert torch.equal(fused[d:2 * d], sd["vilt_encoder.vilt." + q + "key.weight"])
    # in-place edits through torch land in the arena and are noticed (shadow refresh trigger)
    v0 = arena._version_sum
    with torch.no_grad():
        vilt.pooler.dense.bias.add_(1.0)
    arena.sync(allow_cpu=True)
    assert arena._version_sum != v0
    assert torch.equal(arena.theta[off["pooler.dense.bias"]: off["pooler.dense.bias"] + d], vilt.pooler.dense.bias)
    # load_state_dict copies in place: still views
    learner.load_state_dict(vo.synth_state_dict(TINY, ALL_TASKS, seed=6), strict=False)
    assert arena.sync(allow_cpu=True) is False
    # replacing a module (expand / reallocate) forces a rebuild
    learner.vilt_encoder.reallocate_text_image(vilt.embeddings.text_embeddings.position_embeddings.weight.data.clone(), 16, 32)
    assert arena.sync(allow_cpu=True) is True
    assert vilt.embeddings.text_embeddings.position_embeddings.weight.shape[0] == 16


def test_deepcopy_gives_an_independent_model():
    """train_vqa.py:210,242 deep-copies the learner for best-model tracking."""
    learner = _learner()
    learner.vilt_encoder.vilt._arena.sync(allow_cpu=True)
    clone = copy.deepcopy(learner)
    a, b = learner.vilt_encoder.vilt, clone.vilt_encoder.vilt
    assert b._arena is not a._arena and b._arena.owner is b
    b._arena.sync(allow_cpu=True)
    with torch.no_grad():
        a.pooler.dense.weight.zero_()
    assert b.pooler.dense.weight.abs().sum() > 0
    assert set(clone.state_dict().keys()) == set(learner.state_dict().keys())


def test_static_tables_flags_follow_requires_grad():
    from climb_b200 import _lib
    learner = _learner()
    vilt = learner.vilt_encoder.vilt
    vilt._arena.sync(allow_cpu=True)
    st = vilt._static_tables()
    assert st["dims"].pos_grid == 2 and st["dims"].n_modality == 3
    assert all(st["layers"][i].flags == _lib.TRAIN_BASE for i in range(TINY.num_hidden_layers))
    assert st["params"].embed_flags == _lib.TRAIN_BASE and st["params"].adapter_r == 0
    learner.vilt_encoder.freeze_bottom_k_layers(1)
    st = vilt._static_tables()
    assert st["layers"][0].flags == 0 and st["layers"][1].flags == _lib.TRAIN_BASE and st["params"].embed_flags == 0
    learner.add_adapter("vqa", "houlsby")
    learner.train_adapter("vqa")
    vilt._arena.sync(allow_cpu=True)
    st = vilt._static_tables()
    assert st["params"].adapter_r == 8 and st["params"].adapter_act == _lib.EPI_SWISH
    assert all(st["layers"][i].flags == _lib.TRAIN_ADAPTER for i in range(TINY.num_hidden_layers))
    assert st["params"].tail_flags == 0
    assert all(".adapters.vqa." in n for n, _ in st["trainable"])


def test_optimizer_groups_reproduce_the_reference_quirk():
    """Only names containing 'bias' or 'LayerNorm.weight' skip weight decay (src/modeling/vilt.py:209-213):
    layernorm_before/after.weight ARE decayed."""
    learner = _learner()
    opt = learner.create_optimizer({"lr": 1e-4, "weight_decay": 1e-2, "adam_epsilon": 1e-8})
    names = {id(p): n for n, p in learner.named_parameters()}
    decay = [names[id(p)] for p in opt.param_groups[0]["params"]]
    nodecay = [names[id(p)] for p in opt.param_groups[1]["params"]]
    exp_decay, exp_nodecay = vo.weight_decay_groups(list(names.values()))
    assert decay == exp_decay and nodecay == exp_nodecay
    assert "vilt_encoder.vilt.encoder.layer.0.layernorm_before.weight" in decay
    assert "vilt_encoder.vilt.embeddings.text_embeddings.LayerNorm.weight" in nodecay
    assert opt.param_groups[0]["weight_decay"] == 1e-2 and opt.param_groups[1]["weight_decay"] == 0.0
    assert opt.defaults["betas"] == (0.9, 0.98)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: vo.linear_warmup_decay(s, 10, 100))
    assert isinstance(opt, torch.optim.Optimizer) and sched is not None


def test_patch_geometry_from_pixel_mask():
    """pixel_mask -> per-image valid patch rows / cols and the sequence's patch-slot count
    (ViltEmbeddings.visual_embed, modeling_vilt.py:125-129,163-170), host-side, no kernels involved."""
    import torch
    from climb_b200.modeling import B200ViltConfig, B200ViltModel
    from oracle.vilt_oracle import patch_geometry
    m = B200ViltModel(B200ViltConfig(hidden_size=128, num_hidden_layers=1, num_attention_heads=2, intermediate_size=256,
                                     image_size=64, patch_size=32, vocab_size=50, max_position_embeddings=8))
    B, H, W = 3, 96, 160
    mask = torch.zeros(B, H, W, dtype=torch.long)
    sizes = [(96, 160), (64, 96), (32, 160)]
    for b, (h, w) in enumerate(sizes):
        mask[b, :h, :w] = 1
    geom, n = m._patch_geometry(mask, B, H, W, torch.device("cpu"))
    assert geom.tolist() == [[3, 5], [2, 3], [1, 5]] and n == 15 and geom.dtype == torch.int32
    hs, ws = patch_geometry(mask, 32)
    assert hs.tolist() == [3, 2, 1] and ws.tolist() == [5, 3, 5]
    # no padding anywhere -> fixed-resolution path
    assert m._patch_geometry(torch.ones(B, H, W, dtype=torch.long), B, H, W, torch.device("cpu")) == (None, 0)
    assert m._patch_geometry(None, B, H, W, torch.device("cpu")) == (None, 0)
    # a batch whose largest image is smaller than the padded grid keeps only max(h * w) slots
    mask2 = torch.zeros(2, H, W, dtype=torch.long)
    mask2[0, :64, :128] = 1
    mask2[1, :96, :64] = 1
    geom2, n2 = m._patch_geometry(mask2, 2, H, W, torch.device("cpu"))
    assert geom2.tolist() == [[2, 4], [3, 2]] and n2 == 8


def test_viltbert_state_dict_keys_and_registry():
    """B200ViltBertContinualLearner: same checkpoint keys / shapes as src/modeling/viltbert.py's learner
    (viltbert_encoder.vilt.*, viltbert_encoder.bert.*, task_layer.*), BERT without a gradient arena, registry entries."""
    import torch
    from climb_b200 import modeling as M
    bd = vo.BertDims(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, vocab_size=200,
                     max_position_embeddings=16)
    cfg = M.B200ViltConfig(hidden_size=TINY.hidden_size, num_hidden_layers=TINY.num_hidden_layers,
                           num_attention_heads=TINY.num_attention_heads, intermediate_size=TINY.intermediate_size,
                           image_size=TINY.image_size, patch_size=TINY.patch_size, vocab_size=TINY.vocab_size,
                           max_position_embeddings=TINY.max_position_embeddings)
    bcfg = M.B200BertConfig(vocab_size=200, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                            max_position_embeddings=16)
    enc = M.B200ViltBertEncoderWrapper(None, M.B200ViltModel(cfg), M.B200BertModel(bcfg), torch.device("cpu"))
    learner = M.B200ViltBertContinualLearner(ALL_TASKS, enc, 128, vo.TASK_SPECS)
    ref = vo.synth_viltbert_state_dict(TINY, bd, ALL_TASKS, seed=1)
    sd = learner.state_dict()
    assert set(ref) <= set(sd)
    extra = set(sd) - set(ref)
    assert all(k.endswith("position_ids") for k in extra), extra
    for k, v in ref.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    assert learner.get_encoder() is enc and enc.bert._arena.with_grad is False
    assert hasattr(learner, "create_optimizer") and hasattr(learner, "get_active_adapters")
    assert set(M.load_encoder_map) == set(M.create_continual_learner_map) == {"vilt-b200", "viltbert-b200"}
    assert M.model_configs["viltbert-b200"]["encoder_class"] is M.B200ViltBertEncoderWrapper


def test_adapter_handler_surface_matches_the_driver_calls():
    """AdapterHandler as train_upstream_continual_learning.py:167-170, 195-197 uses it (src/cl_algorithms/adapters.py)."""
    import types
    from climb_b200.cl_algorithms import AdapterHandler
    from climb_b200.cl_algorithms.adapters import ADAPTER_MAP, SUPPORTED_ADAPTER_METHODS
    from climb_b200.modeling import AdapterSpec
    assert SUPPORTED_ADAPTER_METHODS == ['vanilla'] and {'houlsby', 'pfeiffer'} <= set(ADAPTER_MAP)
    args = types.SimpleNamespace(adapter_config="houlsby", adapter_reduction_factor=4, ordered_cl_tasks=["vqa", "nlvr2"])
    handler = AdapterHandler("vanilla", args)
    assert handler.adapter_config.reduction_factor == 4 and handler.adapter_config.non_linearity == "swish"
    assert AdapterSpec.from_config("houlsby").reduction_factor == 16, "the override must not leak into the preset"
    learner = _learner()
    handler.add_adapters_to_model(learner)
    r = TINY.hidden_size // 4
    for task in args.ordered_cl_tasks:
        w = dict(learner.named_parameters())[f"vilt_encoder.vilt.encoder.layer.0.output.adapters.{task}.adapter_down.0.weight"]
        assert tuple(w.shape) == (r, TINY.hidden_size)
    handler.activate_adapter_for_training("vqa", learner)
    assert learner.get_active_adapters() == "vqa"
    trainable = {n for n, p in learner.named_parameters() if p.requires_grad}
    assert all(".adapters.vqa." in n or n.startswith("task_layer.") for n in trainable) and any(".adapters.vqa." in n for n in trainable)
    handler.activate_adapter_for_eval("nlvr2", learner)
    assert learner.get_active_adapters() == "nlvr2"
    keep = types.SimpleNamespace(adapter_config="pfeiffer", adapter_reduction_factor=0, ordered_cl_tasks=[])
    assert AdapterHandler("vanilla", keep).adapter_config.reduction_factor == 16      # <= 0 keeps the config's own value
    with pytest.raises(ValueError):
        AdapterHandler("fusion", args)
    with pytest.raises(ValueError):
        AdapterHandler("vanilla", types.SimpleNamespace(adapter_config="compacter", adapter_reduction_factor=0, ordered_cl_tasks=[]))


def test_concat_encodings_checks_shapes():
    from climb_b200.cl_algorithms.experience_replay import concat_encodings
    cur = {"input_ids": torch.zeros(3, 8, dtype=torch.long), "pixel_values": torch.zeros(3, 3, 32, 32), "extra": 1}
    rep = {"input_ids": torch.ones(2, 8, dtype=torch.long), "pixel_values": torch.ones(2, 3, 32, 32)}
    out = concat_encodings(cur, rep)
    assert set(out) == {"input_ids", "pixel_values"} and out["input_ids"].shape == (5, 8) and out["pixel_values"][3:].min() == 1
    rep["pixel_values"] = torch.ones(2, 3, 32, 64)
    with pytest.raises(ValueError):
        concat_encodings(cur, rep)


def _keys_fixture():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")) as f:
        return json.load(f)


def _describe(model):
    sd, enc = model.state_dict(), model.get_encoder().state_dict()
    fmt = lambda d: [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in d.items()]
    return {"state_dict": fmt(sd), "encoder_state_dict": fmt(enc),
            "named_parameters": [[n, bool(p.requires_grad)] for n, p in model.named_parameters()]}


def _same_surface(ours, ref, check_order=True):
    for part in ("state_dict", "encoder_state_dict"):
        a = {k: (tuple(s), d) for k, s, d in ours[part]}
        b = {k: (tuple(s), d) for k, s, d in ref[part]}
        assert set(a) == set(b), (part, sorted(set(a) ^ set(b))[:8])
        assert a == b, (part, [k for k in a if a[k] != b[k]][:8])
    if check_order:
        assert ours["named_parameters"] == ref["named_parameters"]
    else:
        assert dict(map(tuple, ours["named_parameters"])) == dict(map(tuple, ref["named_parameters"]))


def test_checkpoint_surface_equals_the_reference_learners():
    """Every state_dict key / shape / dtype (buffers included), the encoder-only state dict the driver also saves
    (train_upstream_continual_learning.py:264-266) and the named_parameters() order + requires_grad flags of the UNMODIFIED
    reference learners (tests/golden/state_dict_keys.json, oracle/make_golden_keys.py): plain, with Houlsby adapters for two
    tasks after train_adapter('nlvr2'), with a Pfeiffer adapter, and ViLT-BERT."""
    import torch
    from climb_b200 import modeling as M
    g = _keys_fixture()
    learner = _learner()
    _same_surface(_describe(learner), g["vilt"])
    for task in ("vqa", "nlvr2"):
        learner.add_adapter(task, {"reduction_factor": 4, "non_linearity": "swish", "mh_adapter": True, "output_adapter": True})
    learner.train_adapter("nlvr2")
    learner.set_active_adapters("nlvr2")
    # (order included: EWC and the optimizers iterate named_parameters())
    _same_surface(_describe(learner), g["vilt_adapters"], check_order=True)
    learner2 = _learner()
    learner2.add_adapter("snli-ve", {"reduction_factor": 2, "non_linearity": "relu", "mh_adapter": False, "output_adapter": True})
    _same_surface(_describe(learner2), g["vilt_pfeiffer"], check_order=True)
    bd = dict(vocab_size=200, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, max_position_embeddings=16)
    cfg = M.B200ViltConfig(hidden_size=TINY.hidden_size, num_hidden_layers=TINY.num_hidden_layers,
                           num_attention_heads=TINY.num_attention_heads, intermediate_size=TINY.intermediate_size,
                           image_size=TINY.image_size, patch_size=TINY.patch_size, vocab_size=TINY.vocab_size,
                           max_position_embeddings=TINY.max_position_embeddings)
    enc = M.B200ViltBertEncoderWrapper(None, M.B200ViltModel(cfg), M.B200BertModel(M.B200BertConfig(**bd)), torch.device("cpu"))
    vb = M.B200ViltBertContinualLearner(ALL_TASKS, enc, 128, vo.TASK_SPECS)
    _same_surface(_describe(vb), g["viltbert"])
    # and a reference-format checkpoint loads strictly
    ref_sd = {k: torch.zeros(s, dtype=getattr(torch, d)) for k, s, d in g["vilt"]["state_dict"]}
    assert _learner().load_state_dict(ref_sd, strict=True) is not None


def test_ewc_realigns_saved_state_when_the_arena_layout_changes():
    """EWC keeps theta* and F as flat tensors in the arena's layout (cl_algorithms/ewc.py); if parameters are added afterwards
    (an adapter), every saved slice must follow its NAME to the new offset, and the new names must carry no penalty
    (theta* = theta, F = 0) -- the reference keys both dicts by name (src/cl_algorithms/ewc.py:41-43, 82-86)."""
    import types
    from climb_b200.cl_algorithms import EWC
    learner = _learner()
    arena = learner.get_encoder().vilt._arena
    arena.sync(allow_cpu=True)
    ewc = EWC(types.SimpleNamespace(ewc_fisher_sample_percentage=1.0, ewc_loss_weight=100.0))
    assert ewc.do_ewc() is False
    g = torch.Generator().manual_seed(0)
    theta_star = torch.randn(arena.size, generator=g)
    fisher = torch.rand(arena.size, generator=g)
    old_off, old_num = dict(arena.offsets), dict(arena.numels)
    ewc.param_dict["vqa"], ewc.fisher_dict["vqa"], ewc._offsets["vqa"] = theta_star.clone(), fisher.clone(), dict(old_off)
    ewc.fisher_names["vqa"] = list(old_off)
    ewc.task_keys.append("vqa")
    assert ewc.do_ewc() is True
    same_t, same_f = ewc._aligned("vqa", arena)                     # unchanged layout: the saved tensors themselves
    assert same_t is ewc.param_dict["vqa"] and same_f is ewc.fisher_dict["vqa"]
    learner.add_adapter("nlvr2", "houlsby")                          # new parameters in the middle of every layer
    assert arena.sync(allow_cpu=True) is True and arena.offsets != old_off
    new_t, new_f = ewc._aligned("vqa", arena)
    assert new_t.numel() == arena.size and ewc._offsets["vqa"] == arena.offsets
    for n, o in arena.offsets.items():
        k = arena.numels[n]
        if n in old_off:
            assert torch.equal(new_t[o:o + k], theta_star[old_off[n]:old_off[n] + old_num[n]]), n
            assert torch.equal(new_f[o:o + k], fisher[old_off[n]:old_off[n] + old_num[n]]), n
        else:
            assert ".adapters.nlvr2." in n
            assert torch.equal(new_t[o:o + k], arena.theta[o:o + k]) and bool((new_f[o:o + k] == 0).all()), n
    with pytest.raises(TypeError):
        ewc.compute_ewc_loss(types.SimpleNamespace(get_encoder=lambda: types.SimpleNamespace()))


def _local_pretrained_dir(tmp_path):
    """A local directory that ViltProcessor.from_pretrained / ViltConfig.from_pretrained resolve offline (what
    `--pretrained_model_name dandelin/vilt-b32-mlm` resolves through the hub in a CLiMB run), at the tiny geometry."""
    transformers = pytest.importorskip("transformers")
    hub = str(tmp_path / "vilt-tiny-mlm")
    transformers.ViltConfig(hidden_size=TINY.hidden_size, num_hidden_layers=TINY.num_hidden_layers,
                            num_attention_heads=TINY.num_attention_heads, intermediate_size=TINY.intermediate_size,
                            image_size=TINY.image_size, patch_size=TINY.patch_size, vocab_size=TINY.vocab_size,
                            max_position_embeddings=TINY.max_position_embeddings).save_pretrained(hub)
    tok = transformers.BertTokenizerFast(vocab_file=os.path.join(ROOT, "tests", "golden", "tokenizer_vocab.txt"))
    image_cls = getattr(transformers, "ViltImageProcessor", None) or transformers.ViltFeatureExtractor
    transformers.ViltProcessor(image_cls(size={"shortest_edge": 64}), tok).save_pretrained(hub)
    return hub


def test_load_encoder_map_entries_take_the_reference_call(tmp_path):
    """train_language.py:278-279 / train_vision.py:310-311 call `load_encoder_map[name](args.checkpoint_name, device,
    args.pretrained_model_name)` POSITIONALLY (reference signature vilt.py:481, viltbert.py:456): checkpoint = a file written
    by torch.save(encoder.state_dict()) with `vilt.*` (and `bert.*`) keys; an NLVR2-trained checkpoint carries three modality
    rows (vilt.py:98-109, 508-509)."""
    import climb_b200.modeling as M
    hub = _local_pretrained_dir(tmp_path)
    dev = torch.device("cpu")
    learner = _learner()                                         # task list contains nlvr2 -> 3 modality rows
    learner.load_state_dict(vo.synth_state_dict(TINY, ALL_TASKS, seed=8), strict=False)
    ckpt = str(tmp_path / "encoder_after_task_x")                # no 'nlvr2' in the name: the tensor's shape decides
    torch.save(learner.get_encoder().state_dict(), ckpt)
    enc = M.load_encoder_map["vilt-b200"](ckpt, dev, hub)
    assert callable(enc.processor) and enc.processor.tokenizer is not None
    assert enc.vilt.embeddings.token_type_embeddings.weight.shape[0] == 3
    for k, v in learner.get_encoder().state_dict().items():
        assert torch.equal(enc.state_dict()[k], v), k
    # the text half of process_inputs works through the loaded processor (native tokenizer over its vocabulary)
    ids = enc.tokenize(enc.processor.tokenizer, ["is the cat on the mat?", "two dogs"])["input_ids"]
    assert ids.shape[0] == 2 and int(ids[0, 0]) == enc.processor.tokenizer.cls_token_id
    # the tensor's own shape decides: a two-row checkpoint loads as two rows whatever its file name says (the reference
    # expands on the NAME alone and then fails with a size mismatch)
    two = _learner(tasks=["vqa"])
    ckpt2 = str(tmp_path / "encoder_nlvr2_two_rows")
    sd2 = two.get_encoder().state_dict()
    assert sd2["vilt.embeddings.token_type_embeddings.weight"].shape[0] == 2
    torch.save(sd2, ckpt2)
    assert M.load_encoder_map["vilt-b200"](ckpt2, dev, hub).vilt.embeddings.token_type_embeddings.weight.shape[0] == 2
    with pytest.raises(FileNotFoundError):
        M.load_encoder_map["vilt-b200"](str(tmp_path / "missing"), dev, hub)
    # ViLT-BERT: bert.* keys ride in the same checkpoint (viltbert.py:484-485)
    bcfg = M.B200BertConfig(vocab_size=200, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                            max_position_embeddings=16)
    vb = M.B200ViltBertEncoderWrapper(None, learner.get_encoder().vilt, M.B200BertModel(bcfg), dev)
    ckpt3 = str(tmp_path / "viltbert_encoder")
    torch.save(vb.state_dict(), ckpt3)
    got = M.load_encoder_map["viltbert-b200"](ckpt3, dev, hub, bert_config=bcfg)
    for k, v in vb.state_dict().items():
        assert torch.equal(got.state_dict()[k], v), k
    with pytest.raises(RuntimeError):                            # a ViLT-BERT checkpoint through the ViLT loader
        M.load_encoder_map["vilt-b200"](ckpt3, dev, hub)
    # create_continual_learner_map[...](model_name_or_path, ordered_cl_tasks, model_config, task_configs, device): config object
    # in place of the hub name = random init offline
    cl = M.create_continual_learner_map["vilt-b200"](learner.get_encoder().vilt.config, ["vqa"], M.model_configs["vilt-b200"],
                                                     vo.TASK_SPECS, dev)
    assert cl.get_encoder().vilt.grad_sync is None and cl.get_encoder().processor is None


def test_prepare_grads_never_wipes_gradients_it_was_not_asked_about():
    """ParamArena.prepare_grads: an empty `trainable` list (EWC's node when no Fisher-tracked parameter trains) is a no-op,
    and a list that covers only part of the arena zeroes only its own fresh slices while other parameters already hold
    accumulated arena gradients."""
    learner = _learner()
    arena = learner.vilt_encoder.vilt._arena
    arena.sync(allow_cpu=True)
    items = arena.named_items()
    a_name, a_p = items[3]
    b_name, b_p = items[10]
    arena.grad.fill_(7.0)
    arena.publish_grads([(a_name, a_p)])                 # a_p.grad is now the arena slice (value 7)
    arena.prepare_grads([])
    assert float(a_p.grad.flatten()[0]) == 7.0
    arena.prepare_grads([(b_name, b_p)])                 # fresh for b only: a keeps its accumulated gradient
    assert float(a_p.grad.flatten()[0]) == 7.0 and float(arena.grad_view(b_name).abs().max()) == 0.0
    a_p.grad = None
    arena.grad.fill_(7.0)
    arena.prepare_grads([(b_name, b_p)])                 # nothing published anywhere: whole-arena memset
    assert float(arena.grad.abs().max()) == 0.0


def test_max_image_length_draws_the_reference_subsets_on_the_host():
    """B200ViltModel._patch_selection (config.max_image_length > 0, modeling_vilt.py:163-189) against the oracle's restatement
    select_patches, which tests/test_oracle_golden.py pins to the unmodified reference: same seed -> same kept patches, in
    sequence order and in the reference's per-pass order (patch_draw_order); a cap that drops nothing takes the default path."""
    vilt = _learner().get_encoder().vilt
    P = vilt.config.patch_size                      # 16
    H, W = 64, 80                                   # 4 x 5 patch grid
    sizes = [(64, 80), (48, 64), (32, 64), (64, 32), (16, 80), (48, 80)]
    pm = torch.zeros(len(sizes), H, W, dtype=torch.long)
    for k, (h, w) in enumerate(sizes):
        pm[k, :h, :w] = 1
    hs, ws = vo.patch_geometry(pm, P)
    dev = torch.device("cpu")
    for order in (None, [0, 2, 4, 1, 3, 5]):
        torch.manual_seed(7)
        geom, n, sel = vilt._patch_selection(pm, len(sizes), H, W, dev, 9, order)
        torch.manual_seed(7)
        idx = list(range(len(sizes))) if order is None else order
        n_ref, keep = vo.select_patches(hs[idx], ws[idx], (H // P) * (W // P), 9)
        assert n == n_ref == 9 and tuple(sel.shape) == (len(sizes), 9)
        assert geom.tolist() == [[h // P, w // P] for h, w in sizes]
        for j, i in enumerate(idx):
            got = [int(x) for x in sel[i] if x >= 0]
            assert got == keep[j].tolist(), (order, i, got, keep[j].tolist())
            v = (sizes[i][0] // P) * (sizes[i][1] // P)
            assert len(got) == min(v, 9) and all(0 <= x < v for x in got) and len(set(got)) == len(got)
    # a cap at or above the largest image drops nothing: no selection, the default geometry path
    geom, n, sel = vilt._patch_selection(pm, len(sizes), H, W, dev, 20)
    assert sel is None and n == 20
    geom, n, sel = vilt._patch_selection(None, 3, H, W, dev, 20)
    assert geom is None and sel is None
    # no pixel mask at all, cap below the grid: every image draws a subset of the full grid
    torch.manual_seed(3)
    geom, n, sel = vilt._patch_selection(None, 3, H, W, dev, 7)
    assert n == 7 and geom.tolist() == [[4, 5]] * 3 and all(len(set(r.tolist())) == 7 and 0 <= int(r.min()) and int(r.max()) < 20 for r in sel)
    with pytest.raises(ValueError):
        vilt._patch_selection(pm, len(sizes), H, W, dev, 9, [0, 0, 1, 2, 3, 4])
