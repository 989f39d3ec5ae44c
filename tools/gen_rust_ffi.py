"""Prints the `extern "C"` block of the b200pt-sys crate from include/b200pt.h, one Rust declaration
per exported entry point, so that the document cannot drift from the header (tests/test_host_cpu.py checks that every
symbol of the header appears in INTEGRATION.md).  Usage: python tools/gen_rust_ffi.py [--update]   (--update rewrites the block in INTEGRATION.md)"""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TMAP = {"int": "c_int", "int32_t": "i32", "int64_t": "i64", "uint64_t": "u64", "uint32_t": "u32", "uint8_t": "u8", "float": "f32", "double": "f64",
        "char": "c_char", "void": "c_void"}


def rust_type(t):
    t = t.strip()
    ptr = t.count("*")
    const = "const " in t
    base = t.replace("*", "").replace("const", "").strip()
    base = TMAP.get(base, base)
    if ptr == 2:
        return "*mut *mut " + base
    return ("*const " if const else "*mut ") * ptr + base if ptr else base


def prototypes():
    h = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "b200pt.h")).read(), flags=re.S)
    return [" ".join(p.split()) for p in re.findall(r"^(?:int|void|const char\*|void\*|int64_t|int32_t|const b200pt_scene_desc\*)\s+\*?b200pt_\w+\([^;]*\);", h, flags=re.M)]


def declarations():
    out = []
    for p in prototypes():
        ret, star, name, args = re.match(r"(.*?)\s*(\*?)(b200pt_\w+)\((.*)\);", p).groups()
        ret = (ret + star).strip()
        alist = []
        if args.strip() not in ("void", ""):
            for a in args.split(","):
                t, n, arr = re.match(r"(.*?)(\w+)(\[\d*\])?$", a.strip()).groups()
                t = t.strip() + ("*" if arr else "")
                alist.append("%s: %s" % (n + "_" if n in ("type", "ref", "in", "out") else n, rust_type(t)))
        out.append("    pub fn %s(%s)%s;" % (name, ", ".join(alist), "" if ret == "void" else " -> " + rust_type(ret)))
    return out


def update_integration_md():
    """Rewrites the extern block of INTEGRATION.md in place."""
    path = os.path.join(ROOT, "INTEGRATION.md")
    s = open(path).read()
    i0 = s.index('#[link(name = "b200pt")]\nextern "C" {')
    i1 = s.index("```", i0)
    d = declarations()
    s = s[:i0] + '#[link(name = "b200pt")]\nextern "C" {   // generated from include/b200pt.h by tools/gen_rust_ffi.py: all %d entry points\n' % len(d) + "\n".join(d) + "\n}\n" + s[i1:]
    open(path, "w").write(s)


if __name__ == "__main__":
    import sys
    if "--update" in sys.argv:
        update_integration_md()
    else:
        print("\n".join(declarations()))
