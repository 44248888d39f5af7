// hkf_main — headless replacement of the reference's src/main.cpp: hkf_main config.ini [max_frames] [--mesh] [--verbose]
#include <cstdlib>
#include <cstring>
#include <iostream>
#include "MainController.h"

int main(int argc, char** argv)
{
    const char* cfg = argc > 1 ? argv[1] : "config.ini";
    unsigned max_frames = 2000;          // the reference's loop bound (src/MainController.cpp:29-72)
    bool mesh = false, verbose = false;
    for (int i = 2; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--mesh")) mesh = true;
        else if (!std::strcmp(argv[i], "--verbose")) verbose = true;
        else max_frames = (unsigned)std::atoi(argv[i]);
    }
    MainController* mc = MainController::instance();
    if (!mc->init(cfg)) { std::cerr << "init failed: " << mc->lastError() << std::endl; return 1; }
    AppParams::instance()->print();
    const int n = mc->mainLoop(max_frames, mesh, verbose);
    if (n < 0) { std::cerr << mc->lastError() << std::endl; return 2; }
    mc->release();
    return 0;
}
