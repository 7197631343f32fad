"""Host-side helpers mirroring f5_tts_mlx/utils.py (the part on the sample() path): masks, padding,
tokenizers.  Small integer/bool tensors; torch CPU or CUDA, whichever the caller holds."""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F


def exists(v) -> bool:
    return v is not None


def default(v, d):
    return v if exists(v) else d


def lens_to_mask(t: torch.Tensor, length: int | None = None) -> torch.Tensor:
    """utils.py:39-47 — Bool[b, n], mask[b, n] = n < t[b]."""
    if length is None:
        length = int(t.max().item())
    seq = torch.arange(length, device=t.device)
    return seq[None, :] < t[:, None]


def pad_to_length(t: torch.Tensor, length: int, value=0) -> torch.Tensor:
    """utils.py:93-103."""
    if t.ndim not in (1, 2):
        raise ValueError(f"Unsupported padding dims: {t.ndim}")
    seq_len = t.shape[-1]
    if length > seq_len:
        t = F.pad(t, (0, length - seq_len), value=value)
    return t[..., :length]


def pad_sequence(ts: Sequence[torch.Tensor], padding_value=0) -> torch.Tensor:
    """utils.py:106-109."""
    max_len = max(i.shape[-1] for i in ts)
    return torch.stack([pad_to_length(i, max_len, padding_value) for i in ts])


def list_str_to_tensor(text: List[str], padding_value=-1) -> torch.Tensor:
    """utils.py:115-118 — utf-8 byte tokenizer."""
    ts = [torch.tensor([*bytes(t, "UTF-8")], dtype=torch.int32) for t in text]
    return pad_sequence(ts, padding_value=-1)


def list_str_to_idx(text: List[Sequence[str]], vocab_char_map: Dict[str, int], padding_value=-1) -> torch.Tensor:
    """utils.py:124-133 — char/pinyin tokenizer, unknown -> 0, pad -1."""
    ts = [torch.tensor([vocab_char_map.get(c, 0) for c in t], dtype=torch.int32) for t in text]
    return pad_sequence(ts, padding_value=padding_value)


def _segments(text: str):
    """jieba.cut when available (what the reference calls); otherwise, for text without CJK
    characters only, an emulation of its behaviour on such text: runs of [A-Za-z0-9] come out as
    one segment, every other character on its own."""
    try:
        import jieba  # type: ignore
        return list(jieba.cut(text))
    except ImportError:
        if any(len(ch.encode("utf-8")) > 1 and "\u4e00" <= ch <= "\u9fd5" for ch in text):
            raise ImportError("CJK text needs jieba and pypinyin (as in the reference)")
        import re
        return re.findall(r"[A-Za-z0-9]+|.", text, flags=re.S)


def convert_char_to_pinyin(text_list: List[str], polyphone: bool = True) -> List[List[str]]:
    """utils.py:139-173: quote/semicolon normalisation, then per segment: ASCII segments are spelled
    out as characters (a space is inserted before a multi-character segment that directly follows a
    non-space, non-quote character), Chinese segments become space-separated TONE3 pinyin.
    jieba / pypinyin are optional imports (absent from this image)."""
    zh_punc = "。，、；：？！《》【】—…"
    quote_trans = str.maketrans({"“": '"', "”": '"', "‘": "'", "’": "'"})
    oov_trans = str.maketrans({";": ","})
    result: List[List[str]] = []
    for text in text_list:
        chars: List[str] = []
        text = text.translate(quote_trans).translate(oov_trans)
        for seg in _segments(text):
            nbytes = len(seg.encode("utf-8"))
            if nbytes == len(seg):                       # ASCII only
                if chars and nbytes > 1 and chars[-1] not in " :'\"":
                    chars.append(" ")
                chars.extend(seg)
                continue
            from pypinyin import Style, lazy_pinyin  # type: ignore
            if polyphone and nbytes == 3 * len(seg):     # Chinese only
                for syl in lazy_pinyin(seg, style=Style.TONE3, tone_sandhi=True):
                    if syl not in zh_punc:
                        chars.append(" ")
                    chars.append(syl)
            else:                                        # mixed
                for ch in seg:
                    if ord(ch) < 256:
                        chars.append(ch)
                    elif ch in zh_punc:
                        chars.append(ch)
                    else:
                        chars.append(" ")
                        chars.extend(lazy_pinyin(ch, style=Style.TONE3, tone_sandhi=True))
        result.append(chars)
    return result
