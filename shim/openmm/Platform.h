#ifndef OPENMM_PLATFORM_H_
#define OPENMM_PLATFORM_H_
#include "Kernel.h"
#include "KernelFactory.h"
#include "OpenMMException.h"
#include <map>
#include <string>
#include <vector>
namespace OpenMM {
class ContextImpl;
/* Stand-in for OpenMM::Platform: kernel-factory registry per platform + the static platform registry
 * plugins register into (Platform.cpp in the shim holds the statics). */
class OPENMM_EXPORT Platform {
public:
    virtual ~Platform();
    virtual const std::string& getName() const = 0;
    virtual double getSpeed() const { return 1.0; }
    void registerKernelFactory(const std::string& name, KernelFactory* factory);
    Kernel createKernel(const std::string& name, ContextImpl& context) const;
    virtual void contextCreated(ContextImpl&, const std::map<std::string, std::string>&) const {}
    virtual void contextDestroyed(ContextImpl&) const {}
    static void registerPlatform(Platform* platform);
    static int getNumPlatforms();
    static Platform& getPlatform(int index);
    static Platform& getPlatformByName(const std::string& name);
    static void loadPluginLibrary(const std::string& file);
private:
    std::map<std::string, KernelFactory*> kernelFactories;
    static std::vector<Platform*>& getPlatforms();
};
} // namespace OpenMM
#endif
