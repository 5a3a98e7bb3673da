// main.cpp -- `kpeg` CLI drop-in (reference main.cpp:8-140).
//
//   kpeg -h              help text
//   kpeg <file.jpg>      decode on the GPU, write <file>.ppm next to the input, log to kpeg.log + stdout
//   kpeg <in.ppm> <out.jpg>   the reference's encoder entry; its encoder is non-functional
//                        (reference README.md:23) and is out of scope here -- reported as such.
// Exit status follows the reference: EXIT_SUCCESS for any 2- or 3-argument run, EXIT_FAILURE for
// none or too many (main.cpp:54-79).  Extra, non-reference switches come before the file name:
//   --t81        ITU-T T.81 behaviour instead of bit-exact reference parity
//   --device N   CUDA device index
//   --devices L  several devices, e.g. 0-7 or 0,2,3: the restart-interval tiles of the image are spread over them
//   --quiet      log only to kpeg.log
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "Decoder.hpp"
#include "Logger.hpp"

namespace
{
    void printHelp()
    {
        std::cout << "===========================================" << std::endl;
        std::cout << "   K-PEG - Simple JPEG Encoder & Decoder"    << std::endl;
        std::cout << "===========================================" << std::endl;
        std::cout << "Help\n" << std::endl;
        std::cout << "<filename.jpg>                  : Decompress a JPEG image to a PPM image" << std::endl;
        std::cout << "<filename.ppm> <filename.jpg>   : Convert input PNG file to JPEG" << std::endl;
        std::cout << "-h                              : Print this help message and exit" << std::endl;
    }

    // Utility.hpp:16-38: only names whose first ".jpg" is the suffix are accepted (".jpeg" is not).
    bool isValidFilename( const std::string& filename )
    {
        const std::size_t extPos = filename.find( ".jpg" );
        return extPos != std::string::npos && extPos + 4 == filename.size();
    }

    struct Options
    {
        bool parity = true;
        int device = 0;
        std::vector<int> devices;
    };

    // "0-3", "0,2,5", "1": device list of --devices
    std::vector<int> parseDevices( const std::string& s )
    {
        std::vector<int> out;
        std::size_t i = 0;
        while ( i < s.size() )
        {
            std::size_t j = s.find( ',', i );
            if ( j == std::string::npos ) j = s.size();
            const std::string part = s.substr( i, j - i );
            const std::size_t dash = part.find( '-' );
            if ( dash == std::string::npos )
                out.push_back( std::atoi( part.c_str() ) );
            else
                for ( int d = std::atoi( part.substr( 0, dash ).c_str() ); d <= std::atoi( part.substr( dash + 1 ).c_str() ); ++d )
                    out.push_back( d );
            i = j + 1;
        }
        return out;
    }

    void decodeJPEG( const std::string& filename, const Options& opt )
    {
        if ( !isValidFilename( filename ) )
        {
            LOG(kpeg::Logger::Level::ERROR) << "Invalid input file name passed.";
            return;
        }
        kpeg::JPEGDecoder decoder;
        decoder.setParity( opt.parity );
        decoder.setDevice( opt.device );
        if ( !opt.devices.empty() )
            decoder.setDevices( opt.devices );
        decoder.open( filename );
        if ( decoder.decodeImageFile() == kpeg::JPEGDecoder::ResultCode::DECODE_DONE )
            decoder.dumpRawData();
    }
}

int main( int argc, char** argv )
{
    try
    {
        kpeg::Logger::get().openLogFile( "kpeg.log" );
        kpeg::Logger::get().setLevel( kpeg::Logger::Level::DEBUG );

        Options opt;
        std::vector<std::string> args;
        for ( int i = 1; i < argc; ++i )
        {
            const std::string a = argv[i];
            if ( a == "--t81" ) opt.parity = false;
            else if ( a == "--quiet" ) kpeg::Logger::get().setQuiet( true );
            else if ( a == "--device" && i + 1 < argc ) opt.device = std::atoi( argv[++i] );
            else if ( a == "--devices" && i + 1 < argc ) opt.devices = parseDevices( argv[++i] );
            else args.push_back( a );
        }

        LOG(kpeg::Logger::Level::INFO) << "KPEG - Simple JPEG Encoder & Decoder";

        if ( args.empty() )
        {
            LOG(kpeg::Logger::Level::ERROR) << "No arguments provided.";
            return EXIT_FAILURE;
        }
        if ( args.size() == 1 && args[0] == "-h" )
        {
            printHelp();
            return EXIT_SUCCESS;
        }
        if ( args.size() == 1 )
        {
            decodeJPEG( args[0], opt );
            return EXIT_SUCCESS;
        }
        if ( args.size() == 2 )
        {
            LOG(kpeg::Logger::Level::ERROR) << "An error ocurred while encoding. (the encoder is outside this build: "
                                               "the reference's own is non-functional)";
            return EXIT_SUCCESS;
        }
        return EXIT_FAILURE;
    }
    catch ( std::exception& e )
    {
        std::cout << "Exceptions Occurred:-" << std::endl;
        std::cout << "What: " << e.what() << std::endl;
    }
    return EXIT_SUCCESS;
}
