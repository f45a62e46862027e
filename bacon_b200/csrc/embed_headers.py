#!/usr/bin/env python3
"""embed_headers.py HEADER... — prints build/embedded_headers.inc: the device headers as {name, text} initialisers for
rtc.cu (NVRTC gets them as in-memory headers: a user right-hand side compiled at run time needs no source tree)."""
import os
import sys

here = os.path.dirname(os.path.abspath(__file__))


def emit(name, text):
    # split into chunks: string literals have a portable length limit (and MSVC-style 16 KB ones exist)
    out = []
    for i in range(0, len(text), 8000):
        chunk = text[i:i + 8000]
        assert ")BACONHDR" not in chunk
        out.append('R"BACONHDR(' + chunk + ')BACONHDR"')
    print('{"%s",\n%s},' % (name, "\n".join(out)))


for h in sys.argv[1:]:
    emit(h, open(os.path.join(here, h)).read())
abi = open(os.path.join(here, "..", "..", "include", "bacon_ivp.h")).read()
emit("../../include/bacon_ivp.h", abi)
emit("bacon_ivp.h", abi)
# a source written for the registration-macro path (include/bacon_ivp_rhs.cuh) compiles unchanged
emit("bacon_ivp_rhs.cuh", "#pragma once\n#include \"bacon_ivp.h\"\n#define BACON_REGISTER_RHS(RhsType, name)\n")
