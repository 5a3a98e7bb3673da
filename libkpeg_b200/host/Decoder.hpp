// Decoder.hpp -- drop-in for the reference's kpeg::JPEGDecoder (reference include/Decoder.hpp:26-128).
//
// Same public surface and result codes: open() / decodeImageFile() / dumpRawData() / close(),
// ResultCode { SUCCESS, TERMINATE, ERROR, DECODE_INCOMPLETE, DECODE_DONE }.  What is behind it
// differs completely: the file is read once, the container is parsed on the host into a POD plan,
// and the whole hot path -- byteStuffScanData, decodeScanData, MCU construction, IDCT, colour
// conversion, image assembly (reference src/Decoder.cpp:621-855, src/MCU.cpp:64-279,
// src/Image.cpp:20-86) -- runs as CUDA kernels behind the C ABI of include/kpeg_cuda.h.  There is
// no CPU decode path: without a usable GPU decodeImageFile() returns ERROR.
//
// Deliberate differences from the reference, all outside the set of files it decodes correctly:
//   * decoder state is per object (the reference keeps DC predictors in statics and mis-decodes
//     the second image of a process);
//   * DRI/RSTn, APPn, true one-component files and ragged sizes decode;
//   * additions: setParity(), pixels(), width()/height()/components() -- the reference exposes
//     pixels only through the PPM file.
#ifndef KPEG_B200_DECODER_HPP
#define KPEG_B200_DECODER_HPP

#include <cstdint>
#include <iostream>
#include <string>
#include <vector>

#include "kpeg_cuda.h"

namespace kpeg
{
    typedef unsigned char UInt8;

    class JPEGDecoder
    {
        public:
            enum ResultCode
            {
                SUCCESS ,
                TERMINATE ,
                ERROR ,
                DECODE_INCOMPLETE ,
                DECODE_DONE
            };

        public:

            JPEGDecoder();

            // Like the reference (src/Decoder.cpp:18-22) this constructor does NOT open the file.
            JPEGDecoder( const std::string& filename );

            ~JPEGDecoder();

            bool open( const std::string& filename );

            void close();

            ResultCode decodeImageFile();

            // Classify one marker byte (the byte after 0xFF).  The reference's version also parsed the
            // segment from its ifstream (src/Decoder.cpp:53-75); here parsing is kpeg_parse_jfif's
            // job and this only reports what the marker means for the decode.
            ResultCode parseSegmentInfo( const UInt8 byte );

            // Declared but never defined in the reference (include/Decoder.hpp:58); here it lists the
            // markers met while scanning the container.
            void printDetectedSegmentNames();

            bool dumpRawData();

            inline void printCurrPos()
            {
                std::cout << "Current file pos: 0x" << std::hex << m_parsePos << std::dec << std::endl;
            }

        public: // additions

            // true (default): reproduce the reference bit for bit, including its DC-difference quirk
            // (src/MCU.cpp:97-104).  false: ITU-T T.81 behaviour.
            void setParity( bool on ) { m_parity = on; }
            void setDevice( int device ) { m_device = device; m_devices.clear(); }
            // several devices: the restart-interval tiles of the image are spread over them (kpeg_cuda_decode_tiled);
            // an image without usable restart markers is decoded whole by the first one
            void setDevices( const std::vector<int>& devices ) { m_devices = devices; if ( !devices.empty() ) m_device = devices[0]; }

            const std::vector<std::uint8_t>& pixels() const { return m_pixels; }
            unsigned width() const { return m_plan.width; }
            unsigned height() const { return m_plan.height; }
            unsigned components() const { return m_plan.ncomp; }
            const kpeg_stats& stats() const { return m_stats; }

        private:

            std::string m_filename;
            std::vector<std::uint8_t> m_file;
            bool m_opened;
            std::size_t m_parsePos;

            kpeg_plan m_plan;
            kpeg_stats m_stats;
            std::vector<std::uint8_t> m_pixels; // [H][W][ncomp]
            bool m_decoded;
            bool m_parity;
            int m_device;
            std::vector<int> m_devices;
    };
}

#endif
