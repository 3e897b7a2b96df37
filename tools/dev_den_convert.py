"""Development helper (one-off source transformation, kept for reference): rewrites the divisions of the named
device functions in a .cuh file into the branch-free DsDen form (dstar_b200/csrc/dconst.cuh).

    python tools/dev_den_convert.py FILE func1 func2 ...

For each function: adds `template <bool FAST>` and a trailing `bool& bad` parameter, inserts DS_DEN_SCOPE(FAST),
and rewrites `x / tok` into `x / den(tok)` (`den_c(tok)` for literals and physical constants, whose reciprocal
folds at compile time).  Call sites have to be updated by hand.
"""
import re
import sys

CONSTS = {"pi", "amu", "hbar", "clight", "clight2", "boltzmann", "Melectron", "Mneutron", "fourpi", "threepisquare",
          "cm_to_fm", "a_Bohr", "density_n", "Mn_n", "hbarc_n", "electron_compton", "avogadro", "finestructure",
          "onethird", "twothird", "ln10", "Mproton", "arad", "mev_to_ergs", "ergs_to_mev", "fm_to_cm"}


def parse_token(rest):
    """denominator token at the start of `rest`: (expr), ident, ident(...), ident[...], number"""
    if rest.startswith("("):
        depth = 0
        for t, ch in enumerate(rest):
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
                if depth == 0:
                    return rest[:t + 1]
        return None
    m = re.match(r"[A-Za-z_][\w\.:]*(?:<\d+>)?", rest)
    if m:
        tok = m.group(0)
        after = rest[len(tok):]
        if after.startswith("("):
            depth = 0
            for t, ch in enumerate(after):
                if ch == "(":
                    depth += 1
                elif ch == ")":
                    depth -= 1
                    if depth == 0:
                        return tok + after[:t + 1]
            return None
        if after.startswith("["):
            return tok + after[:after.index("]") + 1]
        return tok
    m = re.match(r"[\d\.]+(?:[eE][+-]?\d+)?", rest)
    return m.group(0) if m else None


def convert_body(part, skip_names=()):
    dens = set(re.findall(r"DsDen<FAST> (\w+) =", part)) | set(re.findall(r", (\w+) = den\(", part)) | set(skip_names)
    out, n = [], 0
    for line in part.split("\n"):
        code, sep, com = line.partition("//")
        i, res = 0, ""
        while True:
            j = code.find("/ ", i)
            if j < 0:
                res += code[i:]
                break
            res += code[i:j + 2]
            k = j + 2
            tok = parse_token(code[k:])
            if tok is None or tok in dens or tok.startswith("den(") or tok.startswith("den_c("):
                i = k
                continue
            inner = tok[1:-1] if tok.startswith("(") else tok
            is_const = re.fullmatch(r"[\d\.]+(?:[eE][+-]?\d+)?", tok) or tok in CONSTS or tok.startswith("DS_R4(") \
                or (tok.startswith("dsc::") and tok[5:] in CONSTS)
            res += ("den_c(%s)" if is_const else "den(%s)") % inner
            n += 1
            i = k + len(tok)
        out.append(res + sep + com)
    return "\n".join(out), n


def main():
    path, names = sys.argv[1], sys.argv[2:]
    src = open(path).read()
    for name in names:
        templated = name.startswith("+")   # already `template <bool FAST>` with a `bool& bad` parameter
        name = name.lstrip("+")
        m = re.search(r"\nDS_D (\w[\w:<>]*) %s\(" % re.escape(name), src)
        assert m, name
        a = m.start() + 1
        # end of function: first "\n}\n" after a
        b = src.index("\n}\n", a) + 3
        part = src[a:b]
        hdr_end = part.index(") {\n")
        sig = part[:hdr_end]
        body = part[hdr_end + 4:]
        body, n = convert_body(body)
        if templated:
            part = sig + ") {\n    DS_DEN_SCOPE(FAST);\n" + body
        else:
            part = "template <bool FAST>\n" + sig + ", bool& bad) {\n    DS_DEN_SCOPE(FAST);\n" + body
        src = src[:a] + part + src[b:]
        print(name, "converted", n)
    open(path, "w").write(src)


if __name__ == "__main__":
    main()
