// Decoder.cpp -- kpeg::JPEGDecoder over the CUDA C ABI (see Decoder.hpp).
#include "Decoder.hpp"

#include <cstring>
#include <fstream>

#include "Logger.hpp"

namespace kpeg
{
    namespace
    {
        const char* markerName( unsigned m )
        {
            switch ( m )
            {
                case 0xD8: return "Start of Image (FFD8)";
                case 0xE0: return "JPEG/JFIF Image Marker segment (APP0)";
                case 0xFE: return "Comment(FFFE)";
                case 0xDB: return "Define Quantization Table (FFDB)";
                case 0xC0: return "Start of Frame 0: Baseline DCT (FFC0)";
                case 0xC1: return "Start of Frame 1: Extended Sequential DCT (FFC1), Not supported";
                case 0xC2: return "Start of Frame 2: Progressive DCT (FFC2), Not supported";
                case 0xC4: return "Define Huffman Table (FFC4)";
                case 0xDD: return "Define Restart Interval (FFDD)";
                case 0xDA: return "Start of Scan (FFDA)";
                case 0xD9: return "End of Image (FFD9)";
            }
            if ( m >= 0xE1 && m <= 0xEF ) return "Application segment (APPn)";
            if ( m >= 0xD0 && m <= 0xD7 ) return "Restart marker (RSTn)";
            return "Other segment";
        }
    }

    JPEGDecoder::JPEGDecoder() :
     m_opened( false ), m_parsePos( 0 ), m_plan(), m_stats(), m_decoded( false ), m_parity( true ), m_device( 0 )
    {
        LOG(Logger::Level::INFO) << "Created \'JPEGDecoder object\'.";
    }

    JPEGDecoder::JPEGDecoder( const std::string& /*filename*/ ) :
     m_opened( false ), m_parsePos( 0 ), m_plan(), m_stats(), m_decoded( false ), m_parity( true ), m_device( 0 )
    {
        LOG(Logger::Level::INFO) << "Created \'JPEGDecoder object\'.";
    }

    JPEGDecoder::~JPEGDecoder()
    {
        close();
        LOG(Logger::Level::INFO) << "Destroyed \'JPEGDecoder object\'.";
    }

    bool JPEGDecoder::open( const std::string& filename )
    {
        std::ifstream in( filename, std::ios::in | std::ios::binary );
        if ( !in.is_open() || !in.good() )
        {
            LOG(Logger::Level::ERROR) << "Unable to open image: \'" << filename << "\'";
            return false;
        }
        in.seekg( 0, std::ios::end );
        const std::streamoff n = in.tellg();
        in.seekg( 0, std::ios::beg );
        m_file.resize( n > 0 ? (std::size_t)n : 0 );
        if ( n > 0 )
            in.read( reinterpret_cast<char*>( m_file.data() ), n );
        LOG(Logger::Level::INFO) << "Opened JPEG image: \'" << filename << "\'";
        m_filename = filename;
        m_opened = true;
        m_decoded = false;
        m_parsePos = 0;
        return true;
    }

    void JPEGDecoder::close()
    {
        if ( m_opened )
            LOG(Logger::Level::INFO) << "Closed image file: \'" << m_filename << "\'";
        m_opened = false;
        m_file.clear();
        m_file.shrink_to_fit();
    }

    JPEGDecoder::ResultCode JPEGDecoder::parseSegmentInfo( const UInt8 byte )
    {
        if ( byte == 0x00 || byte == 0xFF )
            return ERROR;
        LOG(Logger::Level::INFO) << "Found segment, " << markerName( byte );
        if ( byte == 0xC1 || byte == 0xC2 )
            return TERMINATE;
        return SUCCESS;
    }

    void JPEGDecoder::printDetectedSegmentNames()
    {
        // Walk the marker segments up to SOS (same traversal kpeg_parse_jfif does).
        std::size_t i = 2;
        if ( m_file.size() < 4 || m_file[0] != 0xFF || m_file[1] != 0xD8 )
            return;
        parseSegmentInfo( 0xD8 );
        while ( i + 4 <= m_file.size() && m_file[i] == 0xFF )
        {
            const unsigned m = m_file[i + 1];
            if ( m == 0xFF ) { ++i; continue; }
            parseSegmentInfo( (UInt8)m );
            if ( m == 0xDA || m == 0xD9 )
                break;
            const std::size_t len = ( (std::size_t)m_file[i + 2] << 8 ) | m_file[i + 3];
            i += 2 + len;
            m_parsePos = i;
        }
    }

    JPEGDecoder::ResultCode JPEGDecoder::decodeImageFile()
    {
        if ( !m_opened || m_file.empty() )
        {
            LOG(Logger::Level::ERROR) << "Unable scan image file: \'" << m_filename << "\'";
            return ResultCode::ERROR;
        }

        LOG(Logger::Level::INFO) << "Started decoding process...";
        printDetectedSegmentNames();

        // one interleaved scan (all the reference decodes, Decoder.cpp:461-530) or one scan per component (T.81 A.2.3)
        kpeg_scan scans[KPEG_MAX_SCANS];
        int nscans = 0;
        const int prc = kpeg_parse_jfif_scans( m_file.data(), m_file.size(), &m_plan, scans, KPEG_MAX_SCANS, &nscans );
        const std::size_t scanOff = nscans ? scans[0].off : 0, scanLen = nscans ? scans[0].len : 0;
        if ( prc == KPEG_ERR_UNSUPPORTED )
        {
            LOG(Logger::Level::INFO) << "Terminated decoding process [NOT-OK].";
            return ResultCode::TERMINATE;
        }
        if ( prc != KPEG_OK )
        {
            LOG(Logger::Level::ERROR) << "[ FATAL ] Invalid JFIF file! Terminating...";
            return ResultCode::ERROR;
        }
        m_plan.flags = m_parity ? KPEG_FLAG_REF_PARITY : 0u;

        m_pixels.assign( (std::size_t)m_plan.width * m_plan.height * m_plan.ncomp, 0 );
        int rc;
        if ( m_devices.size() > 1 && nscans == 1 )
        {
            // restart-interval tiles over several GPUs, every band into its rows of m_pixels (Image.cpp:51-70 placement)
            rc = kpeg_cuda_decode_tiled( m_devices.data(), (int)m_devices.size(), &m_plan, m_file.data() + scanOff, scanLen,
                                         m_pixels.data(), &m_stats );
            if ( rc == KPEG_ERR_CUDA )
                LOG(Logger::Level::ERROR) << "No usable CUDA device: this build has no CPU decode path";
            else if ( rc != KPEG_OK )
                LOG(Logger::Level::ERROR) << "Decode failed: " << kpeg_tiled_last_error();
        }
        else
        {
            // a context borrowed from the process-wide pool: the reference constructs one decoder per file
            // (main.cpp:54-79); streams, pinned bookkeeping and device scratch survive from one file to the next
            kpeg_ctx* ctx = nullptr;
            if ( kpeg_cuda_acquire( m_device, &ctx ) != KPEG_OK )
            {
                LOG(Logger::Level::ERROR) << "No usable CUDA device: this build has no CPU decode path";
                return ResultCode::ERROR;
            }
            rc = kpeg_cuda_decode_scans( ctx, &m_plan, scans, nscans, m_file.data(), m_file.size(), m_pixels.data(), &m_stats );
            if ( rc != KPEG_OK )
                LOG(Logger::Level::ERROR) << "Decode failed: " << kpeg_cuda_last_error( ctx );
            kpeg_cuda_release( m_device, ctx );
        }
        if ( rc == KPEG_ERR_STREAM )
            return ResultCode::DECODE_INCOMPLETE;
        if ( rc != KPEG_OK )
            return ResultCode::ERROR;

        m_decoded = true;
        LOG(Logger::Level::INFO) << "Finished decoding process [OK].";
        return ResultCode::DECODE_DONE;
    }

    bool JPEGDecoder::dumpRawData()
    {
        // Target name: input up to the first ".jpg" (else ".jpeg") + ".ppm" (reference src/Decoder.cpp:77-88)
        std::size_t extPos = m_filename.find( ".jpg" );
        if ( extPos == std::string::npos )
            extPos = m_filename.find( ".jpeg" );
        const std::string target = m_filename.substr( 0, extPos ) + ".ppm";

        if ( !m_decoded )
        {
            LOG(Logger::Level::ERROR) << "Unable to create dump file \'" << target << "\', Invalid pixel pointer";
            return true; // the reference returns true unconditionally (src/Decoder.cpp:87)
        }
        std::ofstream out( target, std::ios::out | std::ios::binary );
        if ( !out.is_open() || !out.good() )
        {
            LOG(Logger::Level::ERROR) << "Unable to create dump file \'" << target << "\'.";
            return true;
        }
        char header[160];
        const int hl = kpeg_ppm_header( m_plan.width, m_plan.height, header, sizeof header );
        out.write( header, hl );
        if ( m_plan.ncomp == 3 )
            out.write( reinterpret_cast<const char*>( m_pixels.data() ), (std::streamsize)m_pixels.size() );
        else
        {
            // P6 has three samples per pixel: a one-component image is written as R = G = B, which is
            // what the reference's colour path yields for Cb = Cr = 128.
            std::vector<std::uint8_t> row( (std::size_t)m_plan.width * 3 );
            for ( unsigned y = 0; y < m_plan.height; ++y )
            {
                const std::uint8_t* src = m_pixels.data() + (std::size_t)y * m_plan.width;
                for ( unsigned x = 0; x < m_plan.width; ++x )
                    row[3 * x] = row[3 * x + 1] = row[3 * x + 2] = src[x];
                out.write( reinterpret_cast<const char*>( row.data() ), (std::streamsize)row.size() );
            }
        }
        LOG(Logger::Level::INFO) << "Raw image data dumped to file: \'" << target << "\'.";
        return true;
    }
}
