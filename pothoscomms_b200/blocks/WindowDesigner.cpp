// /comms/window_designer: host-side window generator with the reference's call surface and
// "tapsChanged" signal (window/WindowDesigner.cpp:60-135); window maths in TapDesign.cpp.
#include <Pothos/Framework.hpp>

#include <string>
#include <vector>

#include "TapDesign.hpp"

class WindowDesigner : public Pothos::Block
{
public:
    static Block *make(void) { return new WindowDesigner(); }

    WindowDesigner(void)
    {
        this->registerCall(this, POTHOS_FCN_TUPLE(WindowDesigner, setWindowType));
        this->registerCall(this, POTHOS_FCN_TUPLE(WindowDesigner, windowType));
        this->registerCall(this, POTHOS_FCN_TUPLE(WindowDesigner, setWindowArgs));
        this->registerCall(this, POTHOS_FCN_TUPLE(WindowDesigner, windowArgs));
        this->registerCall(this, POTHOS_FCN_TUPLE(WindowDesigner, setNumTaps));
        this->registerCall(this, POTHOS_FCN_TUPLE(WindowDesigner, numTaps));
        this->registerSignal("tapsChanged");
    }

    void setWindowType(const std::string &type) { _type = type; this->update(); }
    std::string windowType(void) const { return _type; }
    void setWindowArgs(const std::vector<double> &args) { _args = args; this->update(); }
    std::vector<double> windowArgs(void) const { return _args; }
    void setNumTaps(const size_t num) { _numTaps = num; this->update(); }
    size_t numTaps(void) const { return _numTaps; }
    void activate(void) { this->update(); }

private:
    void update(void)
    {
        if (not this->isActive()) return;
        if (_numTaps == 0) throw Pothos::Exception("WindowDesigner()", "num taps must be positive");   // :125
        try {
            this->emitSignal("tapsChanged", b200c_design::design_window(_type, _numTaps, _args.empty() ? 0.0 : _args.at(0)));
        }
        catch (const std::runtime_error &err) {
            throw Pothos::InvalidArgumentException("WindowDesigner(" + _type + ")", err.what());
        }
    }

    std::string _type = "hann";     // window/WindowDesigner.cpp:61-62
    std::vector<double> _args;
    size_t _numTaps = 51;
};

static Pothos::BlockRegistry registerWindowDesigner("/comms/window_designer", &WindowDesigner::make);
