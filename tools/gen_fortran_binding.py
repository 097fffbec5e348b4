"""Generate fortran/mo_rrnn_c_binding.F90 -- one ISO_C_BINDING interface block per entry point of include/rrnn.h.

    python tools/gen_fortran_binding.py            (rewrites the file)
    python tools/gen_fortran_binding.py --check    (exit 1 if the committed file is stale)

The mapping is mechanical, so that the Fortran side can never drift from the header (tests/test_fortran_cpu.py re-runs it):
  int / float / size_t / long long by value        -> integer(c_int) / real(c_float) / integer(c_size_t) / integer(c_long_long), value
  const char* / char*                               -> character(kind=c_char) :: x(*)
  int* / const int* / long long* / double*          -> integer(c_int) / integer(c_long_long) / real(c_double) :: x(*)   (by reference)
  const rrnn_gas_t*                                 -> type(rrnn_gas_t) :: x(*)
  const rrnn_model_t* const*                        -> type(c_ptr) :: x(*)           (an array of handles)
  T** (handle out-arguments, void**)                -> type(c_ptr) :: x              (by reference)
  every other pointer (handles, float*, void*, unsigned char*: device or host addresses)  -> type(c_ptr), value
  return int / long long / pointer                  -> integer(c_int) / integer(c_long_long) / type(c_ptr)
Host arrays are passed as c_loc(array) (the veneer does); device arrays are c_ptr values anyway.
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rrnn.h")
OUT = os.path.join(ROOT, "fortran", "mo_rrnn_c_binding.F90")


def parse_header(path=HEADER):
    """-> [(name, return C type, [(C type, arg name), ...], preceding comment)]"""
    src = open(path).read()
    # keep the comment that precedes each declaration (it cites the reference interface)
    out = []
    src = src.replace("#define RRNN_API", "#define RRNN_API_")   # (the macro's own definition is not a declaration)
    for m in re.finditer(r"((?:/\*(?:[^*]|\*(?!/))*\*/\s*)*)RRNN_API\s+([^;]+?);", src):
        comment = " ".join(re.sub(r"/\*|\*/|^\s*\*", " ", m.group(1), flags=re.M).split())
        decl = " ".join(m.group(2).split())
        mm = re.match(r"(.+?)\b(rrnn_\w+)\s*\((.*)\)$", decl)
        ret, name, args = mm.group(1).strip(), mm.group(2), mm.group(3).strip()
        alist = []
        if args and args != "void":
            for a in args.split(","):
                a = re.sub(r"/\*.*?\*/", "", a).strip()
                am = re.match(r"(.+?)(\w+)$", a)
                alist.append((am.group(1).strip(), am.group(2)))
        out.append((name, ret, alist, comment))
    return out


def f_arg(ctype, name):
    """-> (Fortran declaration, set of iso_c_binding names to import)"""
    t = ctype.replace("const ", "").replace(" const", "").strip()
    if t == "int":
        return f"integer(c_int), value :: {name}", {"c_int"}
    if t == "float":
        return f"real(c_float), value :: {name}", {"c_float"}
    if t == "size_t":
        return f"integer(c_size_t), value :: {name}", {"c_size_t"}
    if t == "long long":
        return f"integer(c_long_long), value :: {name}", {"c_long_long"}
    if t == "char*":
        return f"character(kind=c_char) :: {name}(*)", {"c_char"}
    if t == "int*":
        return f"integer(c_int) :: {name}(*)", {"c_int"}
    if t == "long long*":
        return f"integer(c_long_long) :: {name}(*)", {"c_long_long"}
    if t == "double*":
        return f"real(c_double) :: {name}(*)", {"c_double"}
    if t == "rrnn_gas_t*":
        return f"type(rrnn_gas_t) :: {name}(*)", {"rrnn_gas_t"}
    if t.endswith("**") or t.endswith("* *"):
        if "rrnn_model_t" in t and ctype.count("const") >= 1 and ctype.strip().startswith("const"):
            return f"type(c_ptr) :: {name}(*)", {"c_ptr"}       # const rrnn_model_t* const*: array of handles
        return f"type(c_ptr) :: {name}", {"c_ptr"}               # handle / pointer out-argument, by reference
    if t.endswith("*"):
        return f"type(c_ptr), value :: {name}", {"c_ptr"}
    raise ValueError(f"unmapped C type {ctype!r} ({name})")


def f_ret(ctype):
    t = ctype.replace("const ", "").strip()
    if t == "int":
        return "integer(c_int)", {"c_int"}
    if t == "long long":
        return "integer(c_long_long)", {"c_long_long"}
    if t.endswith("*"):
        return "type(c_ptr)", {"c_ptr"}
    raise ValueError(f"unmapped C return type {ctype!r}")


FORTRAN_KEYWORDS_RENAME = {"out": "out_h", "value": "val", "name": "name_c", "in": "in_p", "kind": "kind_i", "data": "data_p", "len": "len_i"}


def generate():
    decls = parse_header()
    L = []
    w = L.append
    w("! ISO_C_BINDING interfaces of librrnn_b200.so -- GENERATED from include/rrnn.h by tools/gen_fortran_binding.py; do not edit.")
    w("!")
    w(f"! One interface block per C entry point ({len(decls)} of {len(decls)}).  Scalars by value; handles, device addresses and host")
    w("! arrays as type(c_ptr) values (host arrays: c_loc(a)); out-arguments and small integer / double arrays by reference.")
    w("! The comment above each block is the one the header carries: it cites the reference interface the entry point replaces.")
    w("! NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Fortran compiler, SURVEY.md section 0 F1): `make -C fortran` builds it where")
    w("! gfortran / nvfortran exist; tests/test_fortran_cpu.py parses it (numpy.f2py.crackfortran) and checks every bind(C) name,")
    w("! argument count and argument kind against the header.")
    w("module mo_rrnn_c_binding")
    w("  use, intrinsic :: iso_c_binding")
    w("  implicit none")
    w("  public")
    w("")
    w("  ! rrnn_gas_t (include/rrnn.h): one gas of ty_gas_concs (rrtmgp/mo_gas_concentrations.F90:50-88)")
    w("  type, bind(C) :: rrnn_gas_t")
    w("    character(kind=c_char) :: name(32)")
    w("    type(c_ptr)            :: conc")
    w("    real(c_float)          :: value")
    w("    integer(c_int)         :: ndims")
    w("  end type rrnn_gas_t")
    w("")
    w("  ! activation codes (neural/mod_layer.F90:64-95) and MLP kernel ids")
    w("  integer(c_int), parameter :: RRNN_ACT_LINEAR = 0, RRNN_ACT_SOFTSIGN = 1, RRNN_ACT_RELU = 2, RRNN_ACT_SIGMOID = 3, &")
    w("                               RRNN_ACT_HARD_SIGMOID = 4")
    w("  integer(c_int), parameter :: RRNN_NN_KERNEL_NONE = 0, RRNN_NN_KERNEL_FFMA = 1, RRNN_NN_KERNEL_TCGEN05 = 2")
    w("")
    w("  interface")
    for name, ret, args, comment in decls:
        if comment:
            words, line = comment.split(), "    !"
            for wd in words:
                if len(line) + 1 + len(wd) > 128:
                    w(line); line = "    !"
                line += " " + wd
            w(line)
        rt, imp = f_ret(ret)
        fargs = []
        for ct, an in args:
            an = FORTRAN_KEYWORDS_RENAME.get(an, an)
            d, i = f_arg(ct, an)
            fargs.append((an, d))
            imp |= i
        names = ", ".join(a for a, _ in fargs)
        head = f"    function {name}({names}) bind(C, name=\"{name}\") result(rc)"
        if len(head) > 130:   # continuation lines
            parts, cur = [], f"    function {name}("
            for i, (a, _) in enumerate(fargs):
                piece = a + (", " if i + 1 < len(fargs) else "")
                if len(cur) + len(piece) > 120:
                    parts.append(cur + "&"); cur = "        " + piece
                else:
                    cur += piece
            parts.append(cur + ") &")
            parts.append(f"        bind(C, name=\"{name}\") result(rc)")
            for p_ in parts:
                w(p_)
        else:
            w(head)
        w("      import :: " + ", ".join(sorted(imp)))
        for _, d in fargs:
            w("      " + d)
        w(f"      {rt} :: rc")
        w(f"    end function {name}")
    w("  end interface")
    w("")
    w("contains")
    w("")
    w("  ! the reference's error convention: character(len=128), empty = success (rte/mo_rte_lw.F90:88, 140)")
    w("  function rrnn_error_msg(rc) result(error_msg)")
    w("    integer(c_int), intent(in) :: rc")
    w("    character(len=128)         :: error_msg")
    w("    type(c_ptr) :: p")
    w("    character(kind=c_char), pointer :: s(:)")
    w("    integer :: i")
    w("    error_msg = \"\"")
    w("    if (rc == 0) return")
    w("    p = rrnn_last_error()")
    w("    if (.not. c_associated(p)) then")
    w("      error_msg = \"librrnn_b200: error\"")
    w("      return")
    w("    end if")
    w("    call c_f_pointer(p, s, [128])")
    w("    do i = 1, 128")
    w("      if (s(i) == c_null_char) exit")
    w("      error_msg(i:i) = s(i)")
    w("    end do")
    w("  end function rrnn_error_msg")
    w("")
    w("end module mo_rrnn_c_binding")
    return "\n".join(L) + "\n"


if __name__ == "__main__":
    text = generate()
    if "--check" in sys.argv:
        sys.exit(0 if os.path.exists(OUT) and open(OUT).read() == text else 1)
    with open(OUT, "w") as f:
        f.write(text)
    print(f"wrote {OUT}: {len(parse_header())} interface blocks")
