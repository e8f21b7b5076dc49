"""Host-side tokenizer for the drop-in ``whisper`` package.

Mirrors the interface of reference ``whisper/tokenizer.py:131-395`` (``Tokenizer``,
``get_encoding``, ``get_tokenizer``, ``LANGUAGES``, ``TO_LANGUAGE_CODE``) on top of
``tiktoken``.  The BPE vocabularies are loaded from the repacked binaries under
``assets/`` (see ``tools/pack_vocab.py``).  Only ids and text live here; nothing on
the GPU path depends on this module beyond plain integer ids.
"""
from __future__ import annotations

import gzip
import os
import string
import struct
from dataclasses import dataclass, field
from functools import cached_property, lru_cache
from typing import Dict, List, Optional, Tuple

import tiktoken

# code=name table in Whisper's canonical order (the order fixes the language-token ids).
_LANGUAGE_TABLE = (
    "en=english;zh=chinese;de=german;es=spanish;ru=russian;ko=korean;fr=french;ja=japanese;"
    "pt=portuguese;tr=turkish;pl=polish;ca=catalan;nl=dutch;ar=arabic;sv=swedish;it=italian;"
    "id=indonesian;hi=hindi;fi=finnish;vi=vietnamese;he=hebrew;uk=ukrainian;el=greek;ms=malay;"
    "cs=czech;ro=romanian;da=danish;hu=hungarian;ta=tamil;no=norwegian;th=thai;ur=urdu;"
    "hr=croatian;bg=bulgarian;lt=lithuanian;la=latin;mi=maori;ml=malayalam;cy=welsh;sk=slovak;"
    "te=telugu;fa=persian;lv=latvian;bn=bengali;sr=serbian;az=azerbaijani;sl=slovenian;"
    "kn=kannada;et=estonian;mk=macedonian;br=breton;eu=basque;is=icelandic;hy=armenian;"
    "ne=nepali;mn=mongolian;bs=bosnian;kk=kazakh;sq=albanian;sw=swahili;gl=galician;mr=marathi;"
    "pa=punjabi;si=sinhala;km=khmer;sn=shona;yo=yoruba;so=somali;af=afrikaans;oc=occitan;"
    "ka=georgian;be=belarusian;tg=tajik;sd=sindhi;gu=gujarati;am=amharic;yi=yiddish;lo=lao;"
    "uz=uzbek;fo=faroese;ht=haitian creole;ps=pashto;tk=turkmen;nn=nynorsk;mt=maltese;"
    "sa=sanskrit;lb=luxembourgish;my=myanmar;bo=tibetan;tl=tagalog;mg=malagasy;as=assamese;"
    "tt=tatar;haw=hawaiian;ln=lingala;ha=hausa;ba=bashkir;jw=javanese;su=sundanese;yue=cantonese"
)
LANGUAGES: Dict[str, str] = dict(item.split("=") for item in _LANGUAGE_TABLE.split(";"))

# language name (and a few aliases) -> code
TO_LANGUAGE_CODE: Dict[str, str] = {name: code for code, name in LANGUAGES.items()}
TO_LANGUAGE_CODE.update(
    burmese="my", valencian="ca", flemish="nl", haitian="ht", letzeburgesch="lb", pushto="ps",
    panjabi="pa", moldavian="ro", moldovan="ro", sinhalese="si", castilian="es", mandarin="zh",
)

_SPLIT_PATTERN = r"""'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+"""
_UNSPACED_LANGUAGES = {"zh", "ja", "th", "lo", "my", "yue"}


def _load_ranks(name: str) -> Dict[bytes, int]:
    path = os.path.join(os.path.dirname(__file__), "assets", f"{name}.vocab.gz")
    with gzip.open(path, "rb") as fh:
        blob = fh.read()
    (n,) = struct.unpack_from("<I", blob, 0)
    off, ranks = 4, {}
    for r in range(n):
        (ln,) = struct.unpack_from("<H", blob, off)
        off += 2
        ranks[blob[off:off + ln]] = r
        off += ln
    return ranks


def _special_token_names(num_languages: int) -> List[str]:
    names = ["<|endoftext|>", "<|startoftranscript|>"]
    names += [f"<|{code}|>" for code in list(LANGUAGES)[:num_languages]]
    names += ["<|translate|>", "<|transcribe|>", "<|startoflm|>", "<|startofprev|>",
              "<|nospeech|>", "<|notimestamps|>"]
    names += [f"<|{i * 0.02:.2f}|>" for i in range(1501)]
    return names


@lru_cache(maxsize=None)
def get_encoding(name: str = "gpt2", num_languages: int = 99) -> tiktoken.Encoding:
    """reference tokenizer.py:330-363 - base BPE ranks + Whisper's special tokens appended."""
    ranks = _load_ranks(name)
    base = len(ranks)
    specials = {tok: base + i for i, tok in enumerate(_special_token_names(num_languages))}
    return tiktoken.Encoding(
        name=f"{name}.tiktoken",
        explicit_n_vocab=base + len(specials),
        pat_str=_SPLIT_PATTERN,
        mergeable_ranks=ranks,
        special_tokens=specials,
    )


@dataclass
class Tokenizer:
    """Thin wrapper around ``tiktoken`` with quick access to the special ids
    (reference tokenizer.py:131-327)."""

    encoding: tiktoken.Encoding
    num_languages: int
    language: Optional[str] = None
    task: Optional[str] = None
    sot_sequence: Tuple[int, ...] = ()
    special_tokens: Dict[str, int] = field(default_factory=dict)

    def __post_init__(self):
        enc = self.encoding
        self.special_tokens.update({s: enc.encode_single_token(s) for s in enc.special_tokens_set})
        seq = [self.sot]
        if self.language is not None:
            codes = tuple(LANGUAGES)[: self.num_languages]
            seq.append(self.sot + 1 + codes.index(self.language))
        if self.task is not None:
            seq.append(self.transcribe if self.task == "transcribe" else self.translate)
        self.sot_sequence = tuple(seq)

    # ---- text <-> ids
    def encode(self, text, **kwargs):
        return self.encoding.encode(text, **kwargs)

    def decode(self, token_ids: List[int], **kwargs) -> str:
        return self.encoding.decode([t for t in token_ids if t < self.timestamp_begin], **kwargs)

    def decode_with_timestamps(self, token_ids: List[int], **kwargs) -> str:
        return self.encoding.decode(token_ids, **kwargs)

    # ---- special ids
    def _special(self, name: str) -> int:
        return self.special_tokens[name]

    @cached_property
    def eot(self) -> int:
        return self.encoding.eot_token

    @cached_property
    def transcribe(self) -> int:
        return self._special("<|transcribe|>")

    @cached_property
    def translate(self) -> int:
        return self._special("<|translate|>")

    @cached_property
    def sot(self) -> int:
        return self._special("<|startoftranscript|>")

    @cached_property
    def sot_lm(self) -> int:
        return self._special("<|startoflm|>")

    @cached_property
    def sot_prev(self) -> int:
        return self._special("<|startofprev|>")

    @cached_property
    def no_speech(self) -> int:
        return self._special("<|nospeech|>")

    @cached_property
    def no_timestamps(self) -> int:
        return self._special("<|notimestamps|>")

    @cached_property
    def timestamp_begin(self) -> int:
        return self._special("<|0.00|>")

    @cached_property
    def language_token(self) -> int:
        if self.language is None:
            raise ValueError("This tokenizer does not have language token configured")
        return self.to_language_token(self.language)

    def to_language_token(self, language: str) -> int:
        token = self.special_tokens.get(f"<|{language}|>")
        if token:
            return token
        raise KeyError(f"Language {language} not found in tokenizer.")

    @cached_property
    def all_language_tokens(self) -> Tuple[int, ...]:
        ids = [i for name, i in self.special_tokens.items() if name.strip("<|>") in LANGUAGES]
        return tuple(ids)[: self.num_languages]

    @cached_property
    def all_language_codes(self) -> Tuple[str, ...]:
        return tuple(self.decode([t]).strip("<|>") for t in self.all_language_tokens)

    @cached_property
    def sot_sequence_including_notimestamps(self) -> Tuple[int, ...]:
        return tuple(self.sot_sequence) + (self.no_timestamps,)

    @cached_property
    def non_speech_tokens(self) -> Tuple[int, ...]:
        """Ids suppressed by ``suppress_tokens="-1"``: speaker tags / non-speech annotations
        (brackets, music notes, ...) while keeping ordinary punctuation
        (reference tokenizer.py:237-274)."""
        singles = list('"#()*+/:;<=>@[\\]^_`{|}~「」『』')
        multis = "<< >> <<< >>> -- --- -( -[ (' (\" (( )) ((( ))) [[ ]] {{ }} ♪♪ ♪♪♪".split()
        music = set("♩♪♫♬♭♮♯")  # U+2640..U+267F: suppressing the first token of each is safe
        enc = self.encoding
        out = {enc.encode(" -")[0], enc.encode(" '")[0]}  # no leading hyphen / quote
        for sym in singles + multis + list(music):
            for ids in (enc.encode(sym), enc.encode(" " + sym)):
                if len(ids) == 1 or sym in music:
                    out.add(ids[0])
        return tuple(sorted(out))

    # ---- word splitting (used by long-form transcription)
    def split_to_word_tokens(self, tokens: List[int]):
        if self.language in _UNSPACED_LANGUAGES:
            return self.split_tokens_on_unicode(tokens)
        return self.split_tokens_on_spaces(tokens)

    def split_tokens_on_unicode(self, tokens: List[int]):
        full = self.decode_with_timestamps(tokens)
        bad = "�"
        words, word_tokens, cur, offset = [], [], [], 0
        for tok in tokens:
            cur.append(tok)
            text = self.decode_with_timestamps(cur)
            if bad not in text or full[offset + text.index(bad)] == bad:
                words.append(text)
                word_tokens.append(cur)
                cur = []
                offset += len(text)
        return words, word_tokens

    def split_tokens_on_spaces(self, tokens: List[int]):
        pieces, piece_tokens = self.split_tokens_on_unicode(tokens)
        words, word_tokens = [], []
        for piece, ids in zip(pieces, piece_tokens):
            starts_word = (ids[0] >= self.eot or piece.startswith(" ")
                           or piece.strip() in string.punctuation or not words)
            if starts_word:
                words.append(piece)
                word_tokens.append(ids)
            else:
                words[-1] += piece
                word_tokens[-1].extend(ids)
        return words, word_tokens


@lru_cache(maxsize=None)
def get_tokenizer(multilingual: bool, *, num_languages: int = 99,
                  language: Optional[str] = None, task: Optional[str] = None) -> Tokenizer:
    """reference tokenizer.py:366-395."""
    if language is not None:
        language = language.lower()
        if language not in LANGUAGES:
            if language not in TO_LANGUAGE_CODE:
                raise ValueError(f"Unsupported language: {language}")
            language = TO_LANGUAGE_CODE[language]
    if multilingual:
        name, language, task = "multilingual", language or "en", task or "transcribe"
    else:
        name, language, task = "gpt2", None, None
    return Tokenizer(encoding=get_encoding(name=name, num_languages=num_languages),
                     num_languages=num_languages, language=language, task=task)
