"""CPU tests of the host side of the boundary: tokenizer / detokenizer / sampler against golden vectors produced by the
reference, and the Python surface's streaming / stop-sequence logic (same cases as the reference's tests/test_llm.py)."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

import modelcases

GOLD = np.load(Path(__file__).resolve().parent / "golden" / "host_logic.npz")


@pytest.fixture(scope="module")
def vocabs(lib, tmp_path_factory):
    d = tmp_path_factory.mktemp("vocab")
    out = {}
    for name in ("llama_tiny_q4km", "falcon_tiny_q5km"):
        path, _ = modelcases.build(name, d)
        v = lib.ctb_vocab_load(str(path).encode())
        assert v
        out[name] = v
    yield out
    for v in out.values():
        lib.ctb_vocab_free(v)


@pytest.mark.parametrize("name", ["llama_tiny_q4km", "falcon_tiny_q5km"])
def test_tokenizer_matches_reference(lib, vocabs, name):
    texts = [str(t) for t in GOLD["texts"]]
    buf = (C.c_int * 512)()
    checked = 0
    for i, text in enumerate(texts):
        key = f"{name}_tok_{i}"
        if key not in GOLD:
            continue
        n = lib.ctb_vocab_tokenize(vocabs[name], text.encode(), name.startswith("llama"), buf, 512)
        # the reference's Python layer drops tokens beyond len(text)+1 (its buffer is that small; see llm.py note)
        assert list(buf[:n])[: len(text.encode()) + 1] == GOLD[key].tolist(), text
        checked += 1
    assert checked >= 5


@pytest.mark.parametrize("name", ["llama_tiny_q4km", "falcon_tiny_q5km"])
def test_detokenizer_matches_reference(lib, vocabs, name):
    pieces = GOLD[f"{name}_pieces"]
    assert lib.ctb_vocab_size(vocabs[name]) == len(pieces)
    buf = C.create_string_buffer(256)
    for tok, want in enumerate(pieces):
        n = lib.ctb_vocab_piece(vocabs[name], tok, buf, 256)
        # the reference hands pieces back as C strings, so a piece is cut at its first NUL byte (token <0x00>)
        assert buf.raw[:n].split(b"\0")[0].hex() == str(want), tok
    assert lib.ctb_vocab_piece(vocabs[name], len(pieces) + 5, buf, 256) == 0


@pytest.mark.parametrize("name", ["llama_tiny_q4km", "falcon_tiny_q5km"])
def test_sampler_matches_reference(lib, name):
    logits, picks, settings = GOLD[f"{name}_sample_logits"], GOLD[f"{name}_sample_picks"], GOLD[f"{name}_sample_settings"]
    last = (C.c_int * 3)(5, 6, 7)
    got = []
    for lg in logits:
        lg = np.ascontiguousarray(lg, dtype=np.float32)
        for (k, p, temp, pen, seed) in settings:
            got.append(lib.ctb_sample(lg.ctypes.data_as(C.POINTER(C.c_float)), lg.size, last, 3, int(k), float(p), float(temp), float(pen), int(seed)))
    assert got == picks.tolist()


# ---- Python surface: the reference's own unit test (tests/test_llm.py:4-54 — word-level fake backend, 18 stop cases)
# run against OUR LLM class, plus a character-level variant.
def _mock_llm(words=True):
    from ctransformers_b200.llm import LLM, Config

    class MockLLM(LLM):
        def __init__(self):
            self._config = Config()
            self._llm = None
            self._lib = None

        def tokenize(self, prompt, **kwargs):
            self.pieces = prompt.split(" ") if words else list(prompt)
            return range(len(self.pieces))

        def generate(self, tokens, **kwargs):
            return tokens

        def detokenize(self, tokens, decode=True):
            text = (" " if words else "") + self.pieces[tokens[0]]
            return text if decode else text.encode()

    return MockLLM()


REFERENCE_STOP_CASES = [
    ([], " foo bar baz lorem ipsum\ndolor"), (["dolor "], " foo bar baz lorem ipsum\ndolor"), (["ipsum "], " foo bar baz lorem ipsum\ndolor"),
    (["doloro"], " foo bar baz lorem ipsum\ndolor"), (["ipsumo"], " foo bar baz lorem ipsum\ndolor"), (["dolor"], " foo bar baz lorem ipsum\n"),
    (["ipsum"], " foo bar baz lorem "), (["olor"], " foo bar baz lorem ipsum\nd"), (["olo"], " foo bar baz lorem ipsum\nd"),
    (["psum"], " foo bar baz lorem i"), (["psu"], " foo bar baz lorem i"), (["z lor"], " foo bar ba"), (["rem", "or"], " foo bar baz l"),
    (["foo"], " "), (["f"], " "), ([" "], ""), (["\n"], " foo bar baz lorem ipsum"), (["m\nd"], " foo bar baz lorem ipsu"),
]


@pytest.mark.parametrize("stop,expected", REFERENCE_STOP_CASES)
def test_stop_sequences_reference_cases(stop, expected):
    llm = _mock_llm()
    prompt = "foo bar baz lorem ipsum\ndolor"
    assert llm(prompt, stop=stop) == expected
    assert "".join(llm(prompt, stop=stop, stream=True)) == expected
    if len(stop) == 1:
        assert llm(prompt, stop=stop[0]) == expected


@pytest.mark.parametrize("stop,expected", [(None, "abc xyz abc"), ("x", "abc "), ("c x", "ab"), (["yz", "bc x"], "a"), ("q", "abc xyz abc"),
                                           (["xyz", "y"], "abc x"), ("abc xyz abc", ""), ("abc xyz abcd", "abc xyz abc")])
def test_stop_sequences_char_level(stop, expected):
    llm = _mock_llm(words=False)
    assert llm("abc xyz abc", stop=stop) == expected
    assert "".join(llm("abc xyz abc", stop=stop, stream=True)) == expected


def test_stream_holds_back_partial_stop_prefix():
    llm = _mock_llm(words=False)
    chunks = list(llm("abc xyz", stop=["yzq"], stream=True))
    assert "".join(chunks) == "abc xyz"
    assert chunks[-1].endswith("yz")   # "yz" was held until the end because it could still have become "yzq"


def test_prepare_inputs_reuses_longest_prefix():
    llm = _mock_llm()
    llm._context = [1, 2, 3, 4]
    assert llm.prepare_inputs_for_generation([1, 2, 9, 9]) == [9, 9] and llm._context == [1, 2]
    llm._context = [1, 2, 3]
    assert llm.prepare_inputs_for_generation([1, 2, 3]) == [3] and llm._context == [1, 2]   # keep one token to evaluate
    llm._context = [1, 2, 3]
    assert llm.prepare_inputs_for_generation([7, 8], reset=False) == [7, 8] and llm._context == [1, 2, 3]


def test_auto_config_rejects_unknown_keyword(tmp_path):
    from ctransformers_b200 import AutoConfig
    f = tmp_path / "m.gguf"
    f.write_bytes(b"GGUF")
    with pytest.raises(TypeError):
        AutoConfig.from_pretrained(str(f), not_a_field=1)
    cfg = AutoConfig.from_pretrained(str(f), top_k=3, context_length=128)
    assert cfg.config.top_k == 3 and cfg.config.context_length == 128 and cfg.config.batch_size == 8


def test_model_dir_resolution_and_presets(tmp_path):
    import json
    from ctransformers_b200 import AutoConfig
    from ctransformers_b200.hub import AutoModelForCausalLM
    (tmp_path / "big.gguf").write_bytes(b"GGUF" + b"\0" * 100)
    (tmp_path / "small.gguf").write_bytes(b"GGUF" + b"\0" * 10)
    (tmp_path / "config.json").write_text(json.dumps({"model_type": "llama", "task_specific_params": {"text-generation": {"top_k": 7, "temperature": 0.5}}}))
    assert AutoModelForCausalLM._find_model_file(tmp_path, None).name == "small.gguf"
    assert AutoModelForCausalLM._find_model_file(tmp_path, "big.gguf").name == "big.gguf"
    cfg = AutoConfig.from_pretrained(str(tmp_path))
    assert cfg.model_type == "llama" and cfg.config.top_k == 7 and cfg.config.temperature == 0.5
    with pytest.raises(ValueError):
        AutoConfig.from_pretrained(str(tmp_path / "missing"))
