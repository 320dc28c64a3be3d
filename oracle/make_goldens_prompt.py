"""Generates tests/golden/prompt.json from the UNMODIFIED reference host helpers of row a18:
`tokenizer_image_token` (model/llava/mm_utils.py:19-44), `normalize_cam_params` (datasets/base_contact_dataset.py:37-50) --
both taken out of their files with `ast` and executed as they are (their modules pull in the whole training stack) -- and the
`llava_v1` / `llava_llama_2` conversation templates (model/llava/conversation.py, a self-contained file executed whole) driven
the way run_demo.py:313-323 drives them.  Run in the build container only:  python -m oracle.make_goldens_prompt"""
import ast
import json
import textwrap
import zlib
from pathlib import Path

import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parents[1] / "tests" / "golden" / "prompt.json"


class WordTokenizer:
    """Deterministic stand-in for the LLaMA tokenizer (no sentencepiece model offline): BOS + one id per whitespace word."""
    bos_token_id = 1

    def __call__(self, text):
        ids = [self.bos_token_id] + [3 + zlib.crc32(w.encode()) % 31000 for w in text.split()]
        return type("Enc", (), {"input_ids": ids})()


class NoBosTokenizer(WordTokenizer):
    def __call__(self, text):
        ids = [3 + zlib.crc32(w.encode()) % 31000 for w in text.split()]
        return type("Enc", (), {"input_ids": ids})()


def _extract(path, name, ns):
    src = (REF / path).read_text()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    exec(textwrap.dedent(ast.get_source_segment(src, fn)), ns)
    return ns[name]


QUESTIONS = ["Which body parts are in contact with the chair? Segment these contact areas.",
             "Which part of the mug would a person touch? <image> twice"]
CAMS = [[2., 45., 315., 0., 0.], [2., 315., 315., 0., 0.3], [2., 45., 135., 0., 0.], [2., 315., 135, 0., 0.3], None, [1.5, 330, 225, -0.5, 1.0]]


def main():
    tit = _extract("model/llava/mm_utils.py", "tokenizer_image_token", {"torch": torch, "IMAGE_TOKEN_INDEX": -200})
    ncp = _extract("datasets/base_contact_dataset.py", "normalize_cam_params", {"torch": torch})
    conv_ns = {"__name__": "conversation"}
    exec((REF / "model/llava/conversation.py").read_text(), conv_ns)
    out = {"prompts": [], "cams": [ncp(c).tolist() for c in CAMS]}
    for conv_type in ("llava_v1", "llava_llama_2"):
        for mm in (True, False):
            for q in QUESTIONS:
                conv = conv_ns["conv_templates"][conv_type].copy()
                conv.messages = []
                prompt = "<image>" + "\n" + q                                   # run_demo.py:315
                if mm:
                    prompt = prompt.replace("<image>", "<im_start><image><im_end>")   # :316-320
                conv.append_message(conv.roles[0], prompt)
                conv.append_message(conv.roles[1], "")
                text = conv.get_prompt()
                out["prompts"].append({"conv_type": conv_type, "use_mm_start_end": mm, "question": q, "prompt": text,
                                       "ids_bos": tit(text, WordTokenizer()),
                                       "ids_nobos": tit(text, NoBosTokenizer(), return_tensors="pt").tolist()})
    OUT.write_text(json.dumps(out, indent=0))
    print(len(out["prompts"]), "prompts;", out["prompts"][0]["prompt"][:160].replace("\n", "\\n"), out["prompts"][0]["ids_bos"][:12])


if __name__ == "__main__":
    main()
