"""The LLM example caller (examples/trainer_llm): config schema, wrapper /
artifact helpers and task dispatch on CPU; the whole `decompose_dwain` task on a GPU."""
import copy
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "examples", "trainer_llm")
sys.path.insert(0, EX)
sys.path.insert(0, ROOT)

TINY = {
    "task": "decompose_dwain",
    "decomposed_model_name": "random-init:llama",
    "decomposed_model_revision": "main",
    "decomposed_model_custom_builder_path": None,
    "decomposed_model_custom_builder_config": {
        "vocab_size": 512, "hidden_size": 128, "intermediate_size": 352, "num_hidden_layers": 2,
        "num_attention_heads": 4, "num_key_value_heads": 2, "max_position_embeddings": 256, "seed": 271828},
    "decomposed_model_enable_gradient_checkpointing": False,
    "decomposed_model_dtype": "torch.bfloat16",
    "decomposition_data_name": "synthetic.tokens:2",
    "decomposition_data_separator": "\n\n",
    "decomposition_data_max_length": 128,
    "decomposition_data_batch_size": 2,
    "perplexity_data_name": "synthetic.tokens:3",
    "perplexity_data_separator": "\n\n",
    "perplexity_data_max_length": 128,
    "perplexity_data_batch_size": 2,
    "num_data_steps": 4, "num_metric_steps": 2, "trade_off_factor": 3.0, "reduction_factor": 0.5,
    "max_accepted_ppl_diff": 0.1, "nsr_final_threshold": 1.0, "min_rank": 4, "decompose_in_float64": True,
    "precomputing_covariance_num_splits": 2,
    "finetuning_run": False, "finetuning_use_lora": False, "finetuning_lora_min_rank": 32,
    "finetuning_lr": 0.0001, "finetuning_num_steps": 0, "finetuning_num_last_finetuned_modules": 8,
    "finetuning_use_rank_pattern": False,
    "lm_eval_initial": False, "lm_eval_tasks": None,
    "blacklisted_modules": ["lm_head"],
}

# the field set of the reference's examples_config/decompose_dwain_phi2.yaml, values included
REFERENCE_PHI2 = {
    "ptdeco_trainer_llm_version": "0.1.18", "ptdeco_version": "0.5.7", "task": "decompose_dwain",
    "decomposed_model_name": "microsoft/phi-2", "decomposed_model_revision": "main",
    "decomposed_model_custom_builder_path": None, "decomposed_model_custom_builder_config": None,
    "decomposed_model_enable_gradient_checkpointing": False, "decomposed_model_dtype": "torch.bfloat16",
    "decomposition_data_name": "alpaca.full", "decomposition_data_separator": "\n\n",
    "decomposition_data_max_length": 2048, "decomposition_data_batch_size": 1,
    "perplexity_data_name": "wikitext2.test", "perplexity_data_separator": "\n\n",
    "perplexity_data_max_length": 2048, "perplexity_data_batch_size": 1,
    "num_data_steps": 2048, "num_metric_steps": 32, "trade_off_factor": 3.0, "reduction_factor": 0.5,
    "max_accepted_ppl_diff": 0.1, "nsr_final_threshold": 1.0, "min_rank": 4, "decompose_in_float64": True,
    "precomputing_covariance_num_splits": 4, "finetuning_run": True, "finetuning_use_lora": True,
    "finetuning_lora_min_rank": 32, "finetuning_lr": 0.0001, "finetuning_num_steps": 50,
    "finetuning_num_last_finetuned_modules": 8, "finetuning_use_rank_pattern": False,
    "lm_eval_initial": False,
    "lm_eval_tasks": ["arc_challenge", "arc_easy", "piqa", "hellaswag", "winogrande", "ceval-valid", "cmmlu"],
    "blacklisted_modules": ["lm_head"],
}


def test_config_schema_accepts_reference_fields_and_rejects_unknown_ones():
    import pydantic
    import yaml

    import configurator
    cfg = configurator.DecomposeDWAINConfig(**REFERENCE_PHI2)
    assert cfg.decomposed_model_name == "microsoft/phi-2" and cfg.num_data_steps == 2048
    shipped = yaml.safe_load(open(os.path.join(EX, "examples_config", "decompose_dwain_llama_random.yaml")))
    assert configurator.DecomposeDWAINConfig(**shipped).precomputing_covariance_num_splits == 2
    with pytest.raises(pydantic.ValidationError):
        configurator.DecomposeDWAINConfig(**dict(TINY, no_such_field=1))
    with pytest.raises(pydantic.ValidationError):
        configurator.DecomposeDWAINConfig(**dict(TINY, decomposed_model_dtype="torch.int8"))


def test_dispatch_errors_and_prefix_helpers(tmp_path):
    import collections
    import pathlib

    import dwain_wrapper_module as W
    import run
    with pytest.raises(ValueError, match="Unknown config.task"):
        run.dispatch({"task": "prune"}, pathlib.Path(tmp_path))
    with pytest.raises(ValueError, match="unspecified"):
        run.dispatch({}, pathlib.Path(tmp_path))
    with pytest.raises(ValueError, match="finetune"):
        run.dispatch({"task": "finetune"}, pathlib.Path(tmp_path))
    assert W.add_prefix(["lm_head"]) == ["raw_model.lm_head"]
    assert W.strip_prefix_list(["raw_model.a.b", "c"]) == ["a.b", "c"]
    od = W.strip_prefix_dict(collections.OrderedDict([("raw_model.x.0.weight", 1), ("y", 2)]))
    assert isinstance(od, collections.OrderedDict) and list(od) == ["x.0.weight", "y"]
    W.save_raw_model_decompose_config_and_state_dict(
        pathlib.Path(tmp_path), {"raw_model.m": {"type": "Sequential"}}, {"raw_model.m.0.weight": torch.ones(2, 2)})
    assert json.load(open(tmp_path / "decompose_config.json")) == {"m": {"type": "Sequential"}}
    assert list(torch.load(tmp_path / "decompose_state_dict.pt")) == ["m.0.weight"]


def test_wrapper_loss_builder_and_full_finetune_on_cpu():
    """WrapperModule maps the batch dict to logits, ce_loss is the shifted next-token loss, the
    random-init builder is seeded, and finetune_full trains only the last decomposed modules."""
    import builder
    import datasets_synth
    import dwain_wrapper_module as W
    kw = dict(model_name="random-init:llama", model_revision="main", model_custom_builder_path=None,
              model_custom_builder_config=TINY["decomposed_model_custom_builder_config"],
              enable_gradient_checkpointing=False, dtype=torch.float32)
    m1, tok = builder.make_model_and_tokenizer(**kw)
    m2, _ = builder.make_model_and_tokenizer(**kw)
    assert tok is None and all(torch.equal(a, b) for a, b in zip(m1.state_dict().values(), m2.state_dict().values()))
    builder.validate_module_names(m1, ["lm_head"])
    with pytest.raises(ValueError, match="Unknown module names"):
        builder.validate_module_names(m1, ["no.such.module"])
    data = datasets_synth.SyntheticTokenBatches("synthetic.tokens:5", 512, 16, 2, 4)
    assert len(list(data)) == 4 and torch.equal(data.batch(1)["input_ids"], list(data)[1]["input_ids"])
    with pytest.raises(ValueError, match="only 'synthetic.tokens"):
        datasets_synth.SyntheticTokenBatches("alpaca.full", 512, 16, 2, 4)
    wrapped = W.WrapperModule(m1)
    batch = data.batch(0)
    logits = wrapped(batch)
    assert logits.shape == (2, 16, 512)
    want = torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, 512), batch["labels"][:, 1:].reshape(-1))
    assert torch.allclose(W.ce_loss(batch, logits), want)
    # two "decomposed" modules (two-factor Sequentials); only the last one is fine-tuned
    names = ["raw_model.model.layers.0.mlp.down_proj", "raw_model.model.layers.1.mlp.down_proj"]
    for n in names:
        lin = wrapped.get_submodule(n)
        seq = torch.nn.Sequential(torch.nn.Linear(lin.in_features, 8, bias=False),
                                  torch.nn.Linear(8, lin.out_features, bias=False))
        parent, _, child = n.rpartition(".")
        setattr(wrapped.get_submodule(parent), child, seq)
    before = copy.deepcopy(wrapped.state_dict())
    it = iter(lambda: data.batch(0), None)
    with torch.no_grad():  # the decomposition calls finetune_fn under no_grad
        out = W.finetune_full(model=wrapped, device=torch.device("cpu"), ft_iterator=it,
                              decomposed_modules=names, num_last_modules_to_finetune=1, num_steps=12, lr=1e-2)
    assert out is wrapped and not wrapped.training
    changed = {k for k, v in wrapped.state_dict().items() if not torch.equal(v, before[k])}
    assert changed == {names[1] + ".0.weight", names[1] + ".1.weight"}
    with pytest.raises(RuntimeError, match="peft"):
        W.finetune_lora(model=wrapped)


@pytest.mark.gpu
@pytest.mark.parametrize("finetune", [False, True])
def test_decompose_dwain_task_end_to_end(tmp_path, finetune):
    """The example's whole task on a tiny random-init Llama: artifacts are written under the
    reference's file names with bare-model keys, and load into a freshly built model that then
    reproduces the perplexity the run reported for the decomposed model."""
    import pathlib

    import builder
    import datasets_synth
    import metrics
    import run
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    cfg = dict(TINY)
    if finetune:
        cfg.update(finetuning_run=True, finetuning_num_steps=3, finetuning_num_last_finetuned_modules=2)
    out = pathlib.Path(tmp_path)
    summary = run.dispatch(cfg, out)
    assert summary["modules_decomposed"] >= 1 and summary["mparams_final"] < summary["mparams_initial"]
    assert json.load(open(out / "summary.json"))["perplexity_final"] == summary["perplexity_final"]
    deco = json.load(open(out / "decompose_config.json"))
    assert len(deco) == summary["modules_decomposed"]
    assert all(not k.startswith("raw_model.") and v["type"] == "Sequential" for k, v in deco.items())
    sd = torch.load(out / "decompose_state_dict.pt")
    assert all(not k.startswith("raw_model.") for k in sd)
    assert all(f"{k}.0.weight" in sd and f"{k}.1.weight" in sd for k in deco)
    dev = torch.device("cuda", 0)
    model, _ = builder.make_model_and_tokenizer(
        model_name=cfg["decomposed_model_name"], model_revision="main", model_custom_builder_path=None,
        model_custom_builder_config=cfg["decomposed_model_custom_builder_config"],
        enable_gradient_checkpointing=False, dtype=torch.bfloat16)
    builder.apply_decompose_config_and_state_dict_in_place(
        model=model, decompose_config_path=str(out / "decompose_config.json"),
        state_dict_path=str(out / "decompose_state_dict.pt"), device=dev, dtype=torch.bfloat16)
    loader = datasets_synth.SyntheticTokenBatches(cfg["perplexity_data_name"], 512, 128, 2, 8)
    with torch.no_grad():
        ppl = metrics.calc_perplexity(model, loader, dev, model.config.pad_token_id)
    # same weights; the run evaluated them through the fused two-factor kernel, this model through
    # two F.linear calls: bf16 rounding of the rank-k intermediate differs
    assert abs(ppl - summary["perplexity_final"]) <= 2e-2 * summary["perplexity_final"]


EXV = os.path.join(ROOT, "examples", "trainer_vision")

FALOR_CFG = {
    "task": "decompose_falor", "imagenet_root_dir": "synthetic", "trn_imagenet_classes_fname": None,
    "val_imagenet_classes_fname": None, "batch_size": 4, "normalization": "imagenet", "input_h_w": [64, 64],
    "decompose_model_name": "torchvision.convnext_tiny", "nsr_final_threshold": 0.05, "kl_final_threshold": 0.05,
    "proportion_threshold": 10.0, "num_data_steps": 3, "num_metric_steps": 2, "blacklisted_modules": [],
    "use_float64": False,
}


def _vision_modules():
    """The two example directories reuse module names (builder, configurator, run): load the vision
    ones under their own names."""
    import importlib.util
    mods = {}
    sys.path.insert(0, EXV)
    try:
        for name in ("configurator", "builder", "run_decompose_falor", "run"):
            spec = importlib.util.spec_from_file_location(f"vision_{name}", os.path.join(EXV, f"{name}.py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[spec.name] = mod  # pydantic resolves annotations through sys.modules
            saved = {k: sys.modules.get(k) for k in ("configurator", "builder", "run_decompose_falor")}
            for k in saved:
                if f"vision_{k}" in mods:
                    sys.modules[k] = mods[f"vision_{k}"]
            spec.loader.exec_module(mod)
            for k, v in saved.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
            mods[f"vision_{name}"] = mod
    finally:
        sys.path.remove(EXV)
    return mods


def test_vision_example_config_and_dispatch(tmp_path):
    import pathlib

    import pydantic
    m = _vision_modules()
    cfg = m["vision_configurator"].DecomposeFALORConfig(**FALOR_CFG)
    assert cfg.input_h_w == (64, 64) and cfg.task == "decompose_falor"
    # the reference's examples_config/decompose_falor.yaml fields
    ref = dict(FALOR_CFG, imagenet_root_dir="/nas/datasets/ImageNet", trn_imagenet_classes_fname="train_es.txt",
               val_imagenet_classes_fname="val.txt", batch_size=8, input_h_w=[224, 224],
               decompose_model_name="timm.swinv2_cr_tiny_ns_224.sw_in1k")
    ref.pop("use_float64")
    assert m["vision_configurator"].DecomposeFALORConfig(**ref).use_float64 is False
    with pytest.raises(pydantic.ValidationError):
        m["vision_configurator"].DecomposeFALORConfig(**dict(FALOR_CFG, nope=1))
    with pytest.raises(ValueError, match="Unknown config.task"):
        m["vision_run"].dispatch({"task": "x"}, pathlib.Path(tmp_path))
    with pytest.raises(ValueError, match="timm"):
        m["vision_builder"].make_model("timm.swinv2_cr_tiny_ns_224.sw_in1k")
    it = m["vision_run_decompose_falor"].make_image_iterator(2, (8, 8), "imagenet", 3)
    a, b = next(it), next(it)
    it2 = m["vision_run_decompose_falor"].make_image_iterator(2, (8, 8), "imagenet", 3)
    assert a.shape == (2, 3, 8, 8) and torch.equal(a, next(it2)) and not torch.equal(a, b)


@pytest.mark.gpu
def test_vision_decompose_falor_task_end_to_end(tmp_path):
    """falor through the vision example on a random-init torchvision ConvNeXt-tiny (small synthetic
    images): artifacts under the reference's names load into a freshly built model that then
    produces the decomposed model's logits."""
    import pathlib
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    m = _vision_modules()
    out = pathlib.Path(tmp_path)
    sys.modules["builder"], sys.modules["configurator"] = m["vision_builder"], m["vision_configurator"]
    try:
        summary = m["vision_run_decompose_falor"].main(config_raw=dict(FALOR_CFG), output_path=out)
    finally:
        sys.modules.pop("builder", None)
        sys.modules.pop("configurator", None)
    assert summary["modules_decomposed"] >= 1 and summary["mparams_final"] < summary["mparams_initial"]
    deco = json.load(open(out / "decompose_config.json"))
    assert len(deco) == summary["modules_decomposed"] and all(v["type"] == "Sequential" for v in deco.values())
    dev = torch.device("cuda", 0)
    fresh = m["vision_builder"].make_model(FALOR_CFG["decompose_model_name"])
    m["vision_builder"].apply_decompose_config_and_state_dict_in_place(
        model=fresh, decompose_config_path=str(out / "decompose_config.json"),
        state_dict_path=str(out / "decompose_state_dict.pt"), device=dev)
    assert abs(m["vision_builder"].get_model_stats(fresh)["mparams"] - summary["mparams_final"]) < 1e-9
    x = next(m["vision_run_decompose_falor"].make_image_iterator(4, (64, 64), "imagenet", 1)).to(dev)
    with torch.no_grad():
        y = fresh(x)
    assert torch.isfinite(y).all() and y.shape == (4, 1000)
