"""`AutoConfig` / `AutoModelForCausalLM` for local paths (reference: ctransformers/hub.py:25-268).

The Hugging Face download branch of the reference is out of scope here (no network on the target boxes);
file and directory resolution, `config.json` sampling presets and keyword handling follow the reference.
"""
import json
from dataclasses import dataclass
from pathlib import Path
from typing import Optional

from .llm import CONFIG_FIELDS, Config, LLM

_SAMPLING_KEYS = ("top_k", "top_p", "temperature", "repetition_penalty", "last_n_tokens")


@dataclass
class AutoConfig:
    config: Config
    model_type: Optional[str] = None

    @classmethod
    def from_pretrained(cls, model_path_or_repo_id: str, local_files_only: bool = False, revision: Optional[str] = None, **kwargs) -> "AutoConfig":
        p = Path(model_path_or_repo_id)
        if not (p.is_file() or p.is_dir()):
            raise ValueError(f"Model path '{model_path_or_repo_id}' doesn't exist.")
        auto = cls(config=Config())
        cfg_file = p / "config.json" if p.is_dir() else None
        if cfg_file is not None and cfg_file.is_file():
            meta = json.loads(cfg_file.read_text())
            auto.model_type = meta.get("model_type")
            presets = meta.get("task_specific_params", {}).get("text-generation", {})
            for key in _SAMPLING_KEYS:
                if presets.get(key) is not None:
                    setattr(auto.config, key, presets[key])
        for key, value in kwargs.items():
            if key not in CONFIG_FIELDS:
                raise TypeError(f"'{key}' is an invalid keyword argument for from_pretrained()")
            setattr(auto.config, key, value)
        return auto


class AutoModelForCausalLM:
    @classmethod
    def from_pretrained(cls, model_path_or_repo_id: str, *, model_type: Optional[str] = None, model_file: Optional[str] = None,
                        config: Optional[AutoConfig] = None, lib: Optional[str] = None, local_files_only: bool = False,
                        revision: Optional[str] = None, hf: bool = False, **kwargs) -> LLM:
        if hf:
            raise NotImplementedError("hf=True (transformers adapter) is outside the B200 hot-path build")
        config = config or AutoConfig.from_pretrained(model_path_or_repo_id, local_files_only=local_files_only, revision=revision, **kwargs)
        p = Path(model_path_or_repo_id)
        if p.is_file():
            model_path = p
        else:
            model_path = cls._find_model_file(p, model_file)
        return LLM(model_path=str(model_path), model_type=model_type or config.model_type, config=config.config, lib=lib)

    @staticmethod
    def _find_model_file(directory: Path, model_file: Optional[str]) -> Path:
        """A named file, else the smallest *.gguf / *.bin in the directory (reference: hub.py:233-253)."""
        if model_file:
            f = directory / model_file
            if not f.is_file():
                raise ValueError(f"Model file '{model_file}' not found in '{directory}'")
            return f
        candidates = [f for f in directory.iterdir() if f.is_file() and f.suffix in (".gguf", ".bin")]
        if not candidates:
            raise ValueError(f"No model file found in directory '{directory}'")
        return min(candidates, key=lambda f: f.stat().st_size)
