// AppParamsProducer — config.ini -> AppParams (replaces src/AppParamsProducer.{h,cpp}).  Same sections and keys as
// the reference's config.ini ([switch] [camera] [kinectfusion] [icp] [sdfTracker] [mesh] [IO]); a key that is
// absent keeps the AppParams default instead of crashing (the reference's root config.ini lacks max_iter_nums).
// New optional sections: [synthetic] frames= phase= noise_sigma_mm= ; [switch] use_synthetic_data= stepwise_tracking=.
#pragma once
#include <string>
class AppParamsProducer {
public:
    static bool readParams(const std::string& filename);
    static std::string lastMissingKeys();
};
