"""Width lookup for the text-embedding LLMs.

The reference resolves the hidden size through AutoConfig.from_pretrained
(fusions/load_llm.py:16-35), i.e. a network / HF-cache access.  The values are
fixed properties of the seven aliases it supports (its own _ALIAS table,
fusions/load_llm.py:5-13), so this drop-in answers from a table and only falls
back to transformers for an unknown model id.  Online embedding of raw note
strings (load_llm / embed_notes, fusions/load_llm.py:79-201) is LLM inference,
outside this package's scope: precomputed embeddings are the input."""
from __future__ import annotations

_ALIAS = {
    "GPT2": "openai-community/gpt2",
    "GPT2M": "openai-community/gpt2-medium",
    "GPT2L": "openai-community/gpt2-large",
    "GPT2XL": "openai-community/gpt2-xl",
    "BERT": "google-bert/bert-base-uncased",
    "Llama": "meta-llama/Llama-3.1-8B",
    "DeepSeek": "deepseek-ai/deepseek-llm-7b-base",
}
_D_MODEL = {"GPT2": 768, "GPT2M": 1024, "GPT2L": 1280, "GPT2XL": 1600, "BERT": 768, "Llama": 4096, "DeepSeek": 4096}
_CONTEXT = {"GPT2": 1024, "GPT2M": 1024, "GPT2L": 1024, "GPT2XL": 1024, "BERT": 512, "Llama": 131072, "DeepSeek": 4096}


def register_d_model(alias: str, d_model: int, context_window: int = 1024) -> None:
    """Teach the table a new alias (used by tests and benchmarks for synthetic widths)."""
    _D_MODEL[alias] = int(d_model)
    _CONTEXT[alias] = int(context_window)


def get_d_model(llm_model_fusion: str) -> int:
    if llm_model_fusion in _D_MODEL:
        return _D_MODEL[llm_model_fusion]
    for alias, model_id in _ALIAS.items():
        if model_id == llm_model_fusion:
            return _D_MODEL[alias]
    from transformers import AutoConfig  # unknown id: same lookup as the reference

    cfg = AutoConfig.from_pretrained(_ALIAS.get(llm_model_fusion, llm_model_fusion))
    if hasattr(cfg, "hidden_size"):
        return cfg.hidden_size
    raise AttributeError("Cannot determine hidden size from model/config.")


def get_context_window_size(llm_model_fusion: str, device="cpu") -> int:
    if llm_model_fusion in _CONTEXT:
        return _CONTEXT[llm_model_fusion]
    for alias, model_id in _ALIAS.items():
        if model_id == llm_model_fusion:
            return _CONTEXT[alias]
    raise AttributeError("Cannot determine context window size from model/config.")


def load_llm(*args, **kwargs):
    raise NotImplementedError(
        "Online LLM embedding (use_text_embeddings=False) is outside the B200 fusion hot path; "
        "pass precomputed embeddings (compute_text_embeddings.py) as the reference's main.py does."
    )


def embed_notes(*args, **kwargs):
    load_llm()
