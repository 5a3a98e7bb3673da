#!/usr/bin/env python3
"""TEST INFRASTRUCTURE.  Builds oracle/_ref/kpeg_ref_cuda: the REFERENCE's own kpeg executable (its parser, logger,
Image class and PPM writer, from /root/reference) with its hot path -- decodeScanData() + Image::createImageFromMCUs(),
src/Decoder.cpp:137-138 -- replaced by this repo's C ABI, exactly as INTEGRATION.md section 2 describes.

Nothing of the reference is copied into the repo: its sources are copied into a scratch directory, three edits are
applied there (each checked to have hit), the result is compiled with oracle/hybrid/decode_scan_cuda.cpp and linked
against libkpeg_b200/lib/libkpeg_cuda.so, and the scratch directory is deleted.  Needs /root/reference (build
container only); the binary travels to the GPU box with the snapshot.

usage: build.py [reference_dir]"""
import re
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = ROOT / "oracle" / "_ref" / "kpeg_ref_cuda"
LIBDIR = ROOT / "libkpeg_b200" / "lib"


def edit(path: Path, old: str, new: str, count: int = 1):
    text = path.read_text()
    assert text.count(old) == count, f"{path.name}: expected {count} x {old!r}, found {text.count(old)}"
    path.write_text(text.replace(old, new))


def main():
    if not (REF / "src" / "Decoder.cpp").exists():
        print(f"hybrid: {REF} not present; keeping prebuilt {OUT.name} if any")
        return 0
    if not (LIBDIR / "libkpeg_cuda.so").exists():
        print("hybrid: libkpeg_cuda.so not built yet")
        return 1
    tmp = Path(tempfile.mkdtemp(prefix="kpeg_hybrid_"))
    try:
        for d in ("src", "include"):
            shutil.copytree(REF / d, tmp / d)
        shutil.copy(REF / "main.cpp", tmp / "main.cpp")
        # (1) Decoder.hpp: the byte copy of the scan and the new member
        edit(tmp / "include" / "Decoder.hpp", "std::vector<MCU> m_MCU;",
             "std::vector<MCU> m_MCU;\n            std::vector<UInt8> m_scanBytes;\n            bool decodeScanDataCUDA();")
        # (2) scanImageData (src/Decoder.cpp:558-573): keep the bytes (the '0'/'1' string is not needed any more)
        edit(tmp / "src" / "Decoder.cpp", "m_scanData.append( bits1.to_string() );", "m_scanBytes.push_back( prevByte );")
        edit(tmp / "src" / "Decoder.cpp", "m_scanData.append( bits.to_string() );", "m_scanBytes.push_back( byte );")
        # (3) decodeImageFile (src/Decoder.cpp:137-138): the hot path
        text = (tmp / "src" / "Decoder.cpp").read_text()
        pat = re.compile(r"decodeScanData\(\);\s*m_image\.createImageFromMCUs\( m_MCU \);")
        assert len(pat.findall(text)) == 1
        (tmp / "src" / "Decoder.cpp").write_text(pat.sub("if ( !decodeScanDataCUDA() ) status = ResultCode::ERROR;", text))
        # (4) Image.hpp: the setter next to getPixelPtr()
        edit(tmp / "include" / "Image.hpp", "PixelPtr getPixelPtr();",
             "PixelPtr getPixelPtr();\n            void setPixelPtr( PixelPtr p ) { m_pixelPtr = p; }")
        # quiet log level, as for kpeg_ref_quiet (main.cpp:106)
        edit(tmp / "main.cpp", "setLevel( kpeg::Logger::Level::DEBUG )", "setLevel( kpeg::Logger::Level::ERROR )")
        srcs = ["main.cpp"] + [f"src/{n}.cpp" for n in ("Encoder", "Decoder", "Image", "Logger", "HuffmanTree", "MCU", "Transform")]
        OUT.parent.mkdir(parents=True, exist_ok=True)
        cmd = ["g++", "-O2", "-std=c++14", "-w", "-Iinclude", "-I.", f"-I{ROOT / 'include'}", *srcs, str(HERE / "decode_scan_cuda.cpp"),
               f"-L{LIBDIR}", "-lkpeg_cuda", "-Wl,-rpath,$ORIGIN/../../libkpeg_b200/lib", "-o", str(OUT)]
        subprocess.run(cmd, cwd=tmp, check=True)
        print(f"hybrid: built {OUT}")
        return 0
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    sys.exit(main())
