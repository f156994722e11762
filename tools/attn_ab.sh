for f in 0 1 0 1; do CLIPK_ATTN_FLAGS=$f timeout 100 python tests/attn_prof.py 197 | sed "s/^/flags=$f /"; done
CLIPK_ATTN_V1=1 timeout 100 python tests/attn_prof.py 197 | sed "s/^/v1 /"
