#!/usr/bin/env python3
"""Build oracle/_ref/libref_oracle.so from the reference's OWN sources where they lie
under /root/reference (TEST INFRASTRUCTURE ONLY).

  * device kernels: every kernel of the reference is a C++ raw-string literal
    R"%%( ... )%%" inside a .cc file.  The strings are extracted verbatim into
    oracle/_ref/gen/*.inc (git-ignored) and compiled as host C++ under the macro
    prelude of oracle/ref_prelude.h (KERNEL -> static, GET_GLOBAL_ID() -> the emulated
    work-item, BARRIER_LOCAL -> fiber switch, the -D constants of MakeCompileFlags ->
    run-time globals).  Templated strings get the reference's own TT -> float
    substitution (gen-util.cc:8-14).
  * host code: cuckoo.cc, data.cc, sample.cc, config.cc, types.cc, gen-util.cc are
    (plus random.cc and algorithm/{sum,normalize}.cc
    for link closure) are compiled unmodified against the shape-only stubs in oracle/stubs/.

Nothing is copied into the tracked tree: outputs go to oracle/_ref/ only.  The
library exports the same orc_* C API as liboracle.so (oracle/ammsb_oracle.h), so
the tests can run both through one binding and compare them.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("AMMSB_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
GEN = os.path.join(OUT, "gen")
CXX = "/usr/bin/g++"

RAW = re.compile(r'R"%%\((.*?)\)%%"', re.S)


def strings(relpath):
    return RAW.findall(open(os.path.join(REF, relpath)).read())


def emit(name, text):
    with open(os.path.join(GEN, name), "w") as f:
        f.write("// extracted verbatim from the reference at build time -- not tracked\n")
        f.write(text)
        f.write("\n")


def main():
    if not os.path.isdir(REF):
        print("reference tree not present at %s: keeping any prebuilt oracle/_ref" % REF)
        return 0
    os.makedirs(GEN, exist_ok=True)
    rnd_inc = strings("mcmc/random.cl.inc")
    rnd = strings("mcmc/random.cc")
    ck = strings("mcmc/cuckoo.cc")
    rpm = strings("mcmc/partitioned-alloc.h")
    lrn = strings("mcmc/learner.cc")
    sm = strings("mcmc/algorithm/sum.cc")
    nm = strings("mcmc/algorithm/normalize.cc")
    smp = strings("mcmc/sample.cc")
    phi = strings("mcmc/phi.cc")
    beta = strings("mcmc/beta.cc")
    ppx = strings("mcmc/perplexity.cc")
    assert (len(rnd_inc), len(rnd), len(ck), len(rpm), len(lrn), len(sm), len(nm), len(smp),
            len(phi), len(beta), len(ppx)) == (1, 3, 3, 1, 1, 1, 1, 1, 6, 3, 2), "reference layout changed"
    tt = lambda s: s.replace("TT", "float")  # gen-util.cc:8-14 with type_name<Float>() == "float"
    emit("random_types.inc", rnd[0])
    emit("random_impl.inc", rnd_inc[0])
    emit("random_source.inc", rnd[1])
    emit("gamma.inc", tt(rnd[2]))
    emit("set_types.inc", ck[0])
    emit("set_header.inc", ck[1])
    emit("set_source.inc", ck[2])
    emit("rpm.inc", tt(rpm[0]))
    emit("base_funcs.inc", lrn[0])
    emit("sum.inc", tt(sm[0]))
    emit("normalize.inc", tt(nm[0]))
    emit("sampler.inc", smp[0])
    emit("phi_vec.inc", phi[0])
    emit("phi_thread.inc", phi[1])
    emit("pi_wg.inc", phi[2])
    emit("phi_wg.inc", phi[3])
    emit("beta_base.inc", beta[0])
    emit("beta_thread.inc", beta[1])
    emit("ppx_thread.inc", ppx[0])
    emit("ppx_wg.inc", ppx[1])

    # two builds of the same sources:
    #   libref_oracle.so       the CHECKER: IEEE arithmetic, no contraction (parity tests)
    #   libref_oracle_fast.so  the TIMING build for bench.py's CPU legs: -O3 with fast-math, as the
    #                          reference asks of its own compiler (types.cc:528,532:
    #                          -cl-fast-relaxed-math / -use_fast_math), AVX2 + FMA (not
    #                          -march=native: the library is built here and runs on the GPU box's
    #                          host, whose CPU model is not known at build time)
    builds = (("libref_oracle.so", "chk", ["-O2", "-ffp-contract=off", "-fno-fast-math"]),
              ("libref_oracle_fast.so", "fast", ["-O3", "-ffast-math", "-mavx2", "-mfma"]))
    for lib, tag, opt in builds:
        common = opt + ["-fPIC", "-w", "-fopenmp"]
        objs = []
        for f in ("cuckoo", "data", "sample", "config", "types", "gen-util", "random", "algorithm/sum",
                  "algorithm/normalize"):
            o = os.path.join(OUT, "%s_host_%s.o" % (tag, f.replace("/", "_")))
            subprocess.check_call([CXX, "-std=c++11"] + common + ["-I", os.path.join(HERE, "stubs"), "-I", REF,
                                   "-c", os.path.join(REF, "mcmc", f + ".cc"), "-o", o])
            objs.append(o)
        o = os.path.join(OUT, "%s_ref_api.o" % tag)
        subprocess.check_call([CXX, "-std=gnu++17"] + common + ["-I", os.path.join(HERE, "stubs"), "-I", REF,
                               "-I", HERE, "-I", GEN, "-c", os.path.join(HERE, "ref_api.cc"), "-o", o])
        objs.append(o)
        subprocess.check_call([CXX, "-shared", "-fopenmp", "-o", os.path.join(OUT, lib)] + objs + ["-lm"])
        print("built", os.path.join(OUT, lib))
    return 0


if __name__ == "__main__":
    sys.exit(main())
