// Logger.hpp -- minimal message log of the drop-in CLI / decoder class.
//
// Keeps the look of the reference's log lines ("[ INFO  ][ kpeg:<file>:<line> ] ...", reference
// include/Logger.hpp:39-45,71-80) and its two sinks (kpeg.log + stdout, reference main.cpp:94-104),
// without the reference's pitfalls: the level and the sinks have defaults (the reference leaves
// them uninitialised, Logger.hpp:95-96), and the per-byte / per-MCU DEBUG chatter of the CPU hot
// loops does not exist because those loops run on the GPU.
#ifndef KPEG_B200_LOGGER_HPP
#define KPEG_B200_LOGGER_HPP

#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>

namespace kpeg
{
    class Logger
    {
        public:
            enum Level { ERROR, INFO, DEBUG };

            static Logger& get()
            {
                static Logger instance;
                return instance;
            }

            // Second sink (the CLI passes "kpeg.log"); stdout is always the first one.
            bool openLogFile( const std::string& path )
            {
                m_file.open( path, std::ios::out | std::ios::trunc );
                return m_file.is_open();
            }

            Logger& setLevel( Level level ) { m_level = level; return *this; }
            Level getLevel() const { return m_level; }
            void setQuiet( bool quiet ) { m_stdout = !quiet; }

            void write( Level level, const char* file, int line, const std::string& text )
            {
                if ( level > m_level )
                    return;
                static const char* tag[] = { "[ ERROR ]", "[ INFO  ]", "[ DEBUG ]" };
                const char* base = std::strrchr( file, '/' );
                std::ostringstream ss;
                ss << tag[level] << "[ kpeg:" << ( base ? base + 1 : file ) << ":" << line << " ] " << text << "\n";
                if ( m_stdout )
                    std::cout << ss.str() << std::flush;
                if ( m_file.is_open() )
                    m_file << ss.str() << std::flush;
            }

        private:
            Logger() = default;
            Level m_level = INFO;
            bool m_stdout = true;
            std::ofstream m_file;
    };

    // Collects one line with operator<< and hands it to the logger when it goes out of scope.
    class LogLine
    {
        public:
            LogLine( Logger::Level level, const char* file, int line ) : m_level( level ), m_file( file ), m_line( line ) {}
            ~LogLine() { Logger::get().write( m_level, m_file, m_line, m_ss.str() ); }
            template <typename T> LogLine& operator<<( const T& v ) { m_ss << v; return *this; }
        private:
            Logger::Level m_level;
            const char* m_file;
            int m_line;
            std::ostringstream m_ss;
    };
}

#define LOG(level) ::kpeg::LogLine( level, __FILE__, __LINE__ )

#endif
