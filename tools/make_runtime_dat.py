#!/usr/bin/env python3
"""Generate this repo's trimmed runtime.dat/ from the reference's data tables.

The product keeps XFluids' runtime.dat input FORMAT (species_list.dat per mixture,
thermal_dynamics.dat with NASA-9 blocks; parser: reference src/read_ini/src/thermal.cpp:6-175).
Only the species the five BASELINE configs need are carried (NCOP, H2, O2, N2, AR, Xe); the
coefficient blocks are public NASA Glenn data (NASA/TP-2002-211556) and are emitted verbatim so
that host thermo tables are bit-identical to the reference's.  Run once, in the build container:
    python tools/make_runtime_dat.py /root/reference/runtime.dat runtime.dat
"""
import sys, os, re

SPECIES = ["NCOP", "H2", "O2", "N2", "AR", "Xe"]
MIXTURES = ["NO-COP", "1d-mc-insert-shock-tube", "Inert-SBI", "2d-under-expanded-jet"]


def blocks(path):
    out, cur = {}, None
    for line in open(path):
        if line.startswith("*"):
            cur = line[1:].split()[0] if line[1:].split() else None
            if cur == "END":
                cur = None
                continue
            out[cur] = []
        elif cur is not None:
            out[cur].append(line.rstrip() + "\n")
    return out


def main(src, dst):
    os.makedirs(dst, exist_ok=True)
    b = blocks(os.path.join(src, "thermal_dynamics.dat"))
    with open(os.path.join(dst, "thermal_dynamics.dat"), "w") as f:
        f.write("! xfluids-b200 runtime.dat: NASA-9 thermodynamic coefficients (NASA/TP-2002-211556), XFluids block format:\n")
        f.write("!   star+Name ; 3 x (a1..a7 b1 b2) for 200-1000 K, 1000-6000 K, 6000-20000 K ; molar mass g/mol ; a final star+END line\n")
        f.write("!   Cp/R = a1/T^2 + a2/T + a3 + a4 T + a5 T^2 + a6 T^3 + a7 T^4 ; H/R = -a1/T + a2 ln T + a3 T + ... + b1\n")
        for s in SPECIES:
            f.write("*%s\n" % s)
            f.writelines(b[s])
        f.write("*END\n")
    # Lennard-Jones transport parameters (GRI-Mech transport data as shipped by the reference; parser: thermal.cpp:131-165) for the
    # species carried here, and the Monchick-Mason collision-integral table the transport fits interpolate in (viscfit.cpp:330-350)
    with open(os.path.join(dst, "transport_data.dat"), "w") as f:
        f.write("# xfluids-b200 runtime.dat: Lennard-Jones transport parameters (GRI-Mech 3.0 transport data), XFluids format:\n")
        f.write("# Species  geo  eps/kB(K)  sigma(ang)  mue(Debye)  alpha(ang^3)  Zrot@298K ; the list ends with a star+END line\n")
        seen = set()
        for line in open(os.path.join(src, "transport_data.dat")):
            t = line.split()
            if t and t[0] in SPECIES and t[0] not in seen:
                seen.add(t[0])
                f.write(line.split("!")[0].rstrip() + "\n")
        f.write("*END\n")
    with open(os.path.join(src, "collision_integral.dat")) as fi, open(os.path.join(dst, "collision_integral.dat"), "w") as fo:
        fo.write(fi.read())
    with open(os.path.join(dst, "thermal_dynamics_janaf.dat"), "w") as f:
        f.write("! JANAF/NASA-7 tables are not used (Thermo=1, NASA-9); the parser only needs the terminator\n*END\n")
    for m in MIXTURES:
        os.makedirs(os.path.join(dst, m), exist_ok=True)
        with open(os.path.join(src, m, "species_list.dat")) as fi, open(os.path.join(dst, m, "species_list.dat"), "w") as fo:
            fo.write(fi.read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
