#!/usr/bin/env python
"""Generates bindings/rust/src/lib.rs (the raw Rust FFI binding a Molchanica / `dynamics` maintainer would add) from
include/molchanica_md.h, so that the binding cannot drift from the header.  tests/test_abi.py re-runs it and compares."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TMAP = {'int': 'c_int', 'double': 'f64', 'float': 'f32', 'int64_t': 'i64', 'int32_t': 'i32', 'uint8_t': 'u8', 'uint16_t': 'u16',
        'uint64_t': 'u64', 'char': 'c_char', 'mc_ctx': 'McCtx', 'mc_float4': 'McFloat4', 'mc_energy': 'McEnergy',
        'mc_stats': 'McStats', 'void': 'c_void'}

HEAD = """//! Raw FFI binding of `libmolchanica_md.so` (include/molchanica_md.h) for the reference's host side: what
//! `dynamics` / `src/md` would link in place of the PTX module of build.rs:10-16 (INTEGRATION.md says where each call
//! goes).  GENERATED from the header by tools/gen_rust_binding.py -- do not edit by hand; tests/test_abi.py keeps it
//! in step with the header.  There is no Rust toolchain in the build image: this file has not been compiled.
#![allow(non_camel_case_types, clippy::too_many_arguments)]
use std::os::raw::{c_char, c_int};

"""

FIXED = """
#[repr(C)]
pub struct McCtx {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct McFloat4 {
    pub x: f32,
    pub y: f32,
    pub z: f32,
    pub w: f32,
}

"""

TAIL = """}

/// Maps a status code to the error text `build_dynamics` propagates as `ParamError` (reference src/md/mod.rs:651).
pub fn check(ctx: *const McCtx, rc: c_int) -> Result<(), String> {
    if rc >= MC_OK {
        return Ok(()); // positive codes are warnings (MC_W_STALE_LIST): the call completed, mc_last_error has the text
    }
    let msg = unsafe { std::ffi::CStr::from_ptr(mc_last_error(ctx)) };
    Err(format!("molchanica_md error {}: {}", rc, msg.to_string_lossy()))
}
"""


def _conv(t):
    t = t.strip()
    const = 'const' in t
    t = t.replace('const', '').strip()
    stars = t.count('*')
    base = TMAP[t.replace('*', '').strip()]
    for _ in range(stars):
        base = ('*const ' if const else '*mut ') + base
    return base


def _struct(text, cname, rname):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    body = re.search(r"typedef struct \{([^{}]*)\}\s*" + cname + r"\s*;", text, re.S).group(1)
    fields = []
    for decl in body.split(';'):
        decl = ' '.join(decl.split())
        if not decl:
            continue
        ty, names = decl.split(' ', 1)
        for nm in names.split(','):
            nm = nm.strip()
            arr = re.match(r"(\w+)\[(\d+)\]", nm)
            fields.append(f"    pub {arr.group(1)}: [{TMAP[ty]}; {arr.group(2)}]," if arr else f"    pub {nm}: {TMAP[ty]},")
    return "#[repr(C)]\n#[derive(Clone, Copy, Default, Debug)]\npub struct " + rname + " {\n" + "\n".join(fields) + "\n}\n"


def generate():
    hdr = open(os.path.join(ROOT, "include", "molchanica_md.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    consts = []
    for name, val in re.findall(r"#define (MC_[A-Z0-9_]+) \(?(-?\d+)u?\)?", txt):
        consts.append(f"pub const {name}: {'u8' if name == 'MC_FLAG_STATIC' else 'c_int'} = {val};")
    protos = re.findall(r"\n\s*((?:const\s+)?(?:int|double|char)\s*\*?\s*mc_[a-z0-9_]+\s*\([^;]*?\))\s*;", txt)
    fns = []
    for p in protos:
        p = ' '.join(p.split())
        m = re.match(r"(.*?)(mc_[a-z0-9_]+)\s*\((.*)\)$", p)
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        rargs = []
        if args != 'void':
            for a in args.split(','):
                a = a.strip()
                arr = re.search(r"\[(\d*)\]$", a)
                if arr:
                    a = a[:arr.start()].strip()
                mm = re.match(r"(.*?)([A-Za-z_][A-Za-z0-9_]*)$", a)
                ty, nm = mm.group(1).strip(), mm.group(2)
                if arr:
                    ty += ' *'
                if nm in ('type', 'in', 'ref', 'box'):
                    nm += '_'
                rargs.append(f"{nm}: {_conv(ty)}")
        fns.append(f"    pub fn {name}({', '.join(rargs)}) -> {_conv(ret)};")
    return (HEAD + "\n".join(consts) + "\n" + FIXED + _struct(hdr, "mc_energy", "McEnergy") + "\n" + _struct(hdr, "mc_stats", "McStats") +
            '\n#[link(name = "molchanica_md")]\nextern "C" {\n' + "\n".join(fns) + "\n" + TAIL)


if __name__ == "__main__":
    out = os.path.join(ROOT, "bindings", "rust", "src", "lib.rs")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    open(out, "w").write(generate())
    print("wrote", out)
