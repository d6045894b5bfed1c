"""ORACLE BUILD HELPER -- TEST INFRASTRUCTURE ONLY.

Cuts whole FUNCTION DEFINITIONS out of the reference's own source files, where they lie under /root/reference, into a
generated include file under oracle/_ref/ (git-ignored; deleted again by oracle/Makefile once the object is built), so
that reference functions living in translation units that cannot be compiled here (Frame.cc, MapPoint.cc, ORBmatcher.cc
pull in Eigen / Sophus / g2o / ROS) can still be compiled UNMODIFIED -- character for character -- over the small class
stand-ins of oracle/ref_frame_shim.cpp.  No reference text is stored in this repository: a function is located by the
first line of its definition and ends at the first closing brace in column 0.

    python extract_ref.py <out.inc> <source file>::<first line prefix> ...
"""
import sys


def cut(path, prefix):
    lines = open(path, errors="replace").read().split("\n")
    if prefix.startswith("="):                     # "=<prefix>": that single line only (a #define, a constant)
        for line in lines:
            if line.startswith(prefix[1:]):
                return line + "\n"
        raise SystemExit("extract_ref: '%s' not found in %s" % (prefix, path))
    for i, line in enumerate(lines):
        if line.startswith(prefix):
            out = []
            for l in lines[i:]:
                out.append(l)
                if l.rstrip() == "}":
                    return "\n".join(out) + "\n"
            break
    raise SystemExit("extract_ref: '%s' not found in %s" % (prefix, path))


def main():
    out = sys.argv[1]
    parts = []
    for spec in sys.argv[2:]:
        path, prefix = spec.split("::", 1)
        parts.append("// ---- %s :: %s\n" % (path, prefix) + cut(path, prefix))
    with open(out, "w") as f:
        f.write("\n".join(parts))


if __name__ == "__main__":
    main()
