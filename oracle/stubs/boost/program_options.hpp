// Stand-in for <boost/program_options.hpp>, which this image lacks.
// The reference's util/util.hpp:14,21 includes it only to declare a namespace alias
// (`namespace po = boost::program_options;`); no boost symbol is used.  This lets the
// UNMODIFIED reference headers compile from where they lie under /root/reference.
#pragma once
namespace boost { namespace program_options {} }
