#!/usr/bin/env python
"""Per-code-point tables (BMP) for the native WordPiece tokenizer (easynlp_b200/csrc/wordpiece.cu), generated from the SAME Python
`unicodedata` / `str` predicates the reference's BertTokenizer evaluates (modelzoo/models/bert/tokenization_bert.py: _is_whitespace,
_is_control, _is_punctuation, str.lower + NFD + strip Mn, str.split) -- so the C++ path is exact on the BMP by construction.

    python tools/gen_unicode_table.py out.bin

Layout (little endian): magic 'CLPKUNI1', u32 pool_len, then for cp in [0, 65536): u8 flags, u8 map_len (255 = identity), u32 map_off;
then the pool of u32 code points.  flags: 1 clean->space, 2 removed by clean, 4 punctuation, 8 str.isspace (token split), 16 unsupported
(context-sensitive lower-casing: the caller falls back to the Python tokenizer)."""
import struct
import sys
import unicodedata


def main(out):
    flags = bytearray(65536); mlen = bytearray(65536); moff = [0] * 65536
    pool = []
    for cp in range(65536):
        ch = chr(cp)
        f = 0
        if 0xD800 <= cp <= 0xDFFF:          # surrogates never appear in valid UTF-8
            flags[cp] = 16; mlen[cp] = 255
            continue
        cat = unicodedata.category(ch)
        is_ws = ch in (" ", "\t", "\n", "\r") or cat == "Zs"
        is_ctrl = (ch not in ("\t", "\n", "\r")) and cat.startswith("C")
        if cp == 0 or cp == 0xFFFD or is_ctrl:
            f |= 2
        elif is_ws:
            f |= 1
        if (33 <= cp <= 47) or (58 <= cp <= 64) or (91 <= cp <= 96) or (123 <= cp <= 126) or cat.startswith("P"):
            f |= 4
        if ch.isspace():
            f |= 8
        if cp == 0x03A3:                     # capital sigma: str.lower() applies the Final_Sigma rule
            f |= 16
        low = ch.lower()
        res = [c for c in unicodedata.normalize("NFD", low) if unicodedata.category(c) != "Mn"]
        if len(res) == 1 and res[0] == ch:
            mlen[cp] = 255
        else:
            if len(res) > 250 or any(ord(c) > 0x10FFFF for c in res):
                f |= 16; mlen[cp] = 255
            else:
                mlen[cp] = len(res); moff[cp] = len(pool); pool.extend(ord(c) for c in res)
        flags[cp] = f
    with open(out, "wb") as fh:
        fh.write(b"CLPKUNI1")
        fh.write(struct.pack("<I", len(pool)))
        for cp in range(65536):
            fh.write(struct.pack("<BBI", flags[cp], mlen[cp], moff[cp]))
        fh.write(struct.pack(f"<{len(pool)}I", *pool))
    return len(pool)


if __name__ == "__main__":
    print("pool", main(sys.argv[1]))
