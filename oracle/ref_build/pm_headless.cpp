// Build shim (test infrastructure): a headless stand-in for the reference's
// Win32-only PrintMachine.cpp.  It defines exactly the statics the reference's own
// PrintMachine.h declares (the header itself is taken from the reference, unmodified),
// with the console I/O removed: the frame simply lands in the back buffer.
#include "pch.h"
#include "PrintMachine.h"
#include "Timer.h"

std::mutex PrintMachine::m_Lock;
int PrintMachine::m_renderingFps = 60;
int PrintMachine::m_printingFps = 60;
std::unique_ptr<Time> PrintMachine::m_timer = nullptr;
int PrintMachine::m_printingFpsCounter = 0;
float PrintMachine::m_printingFpsTimer = 0.0f;
size_t PrintMachine::currentWidth = 0;
size_t PrintMachine::currentHeight = 0;
size_t PrintMachine::m_maxSize = 0;
bool PrintMachine::m_running = true;
bool PrintMachine::m_terminateThread = false;
std::unique_ptr<char[]> PrintMachine::m_printBuffer = nullptr;
std::unique_ptr<char[]> PrintMachine::m_backBuffer = nullptr;
size_t PrintMachine::m_printSize = 0;
size_t PrintMachine::m_backBufferPrintSize = 0;
std::string PrintMachine::m_debugInfo = "";
HANDLE PrintMachine::m_inputHandle = nullptr;
HANDLE PrintMachine::m_outputHandle = nullptr;
std::thread PrintMachine::m_printThread;
std::mutex PrintMachine::m_backBufferMutex;
bool PrintMachine::m_bShouldSwapBuffer = false;

void PrintMachine::Start(const size_t x, const size_t y)
{
    currentWidth = x;
    currentHeight = y;
    m_maxSize = m_charsPerPixel * currentWidth * currentHeight;
    m_printBuffer.reset();
    m_backBuffer = std::make_unique<char[]>(m_maxSize);
    m_printSize = m_maxSize;
}
void PrintMachine::CleanUp() { m_backBuffer.reset(); m_printBuffer.reset(); }
bool PrintMachine::CheckIfRunning() { return m_running; }
void PrintMachine::SetDebugInfo(const std::string& s) { m_debugInfo = s; }
void PrintMachine::TerminateThread() { m_terminateThread = true; }
bool PrintMachine::Print() { return true; }
void PrintMachine::UpdateRenderingFPS(const int fps) { m_renderingFps = fps; }
bool PrintMachine::ChangeSize(const size_t x, const size_t y) { currentWidth = x; currentHeight = y; return true; }
const std::mutex* PrintMachine::GetBackBufferMutex() { return &m_backBufferMutex; }
const char* PrintMachine::GetBackBuffer() { return m_backBuffer.get(); }
void PrintMachine::SetDataInBackBuffer(const char* data, const size_t size)
{
    std::lock_guard<std::mutex> g(m_backBufferMutex);
    memcpy(m_backBuffer.get(), data, size);
    m_bShouldSwapBuffer = true;
    m_backBufferPrintSize = size;
}
size_t PrintMachine::GetWidth() { return currentWidth; }
size_t PrintMachine::GetHeight() { return currentHeight; }
size_t PrintMachine::GetMaxSize() { return m_maxSize; }
HANDLE PrintMachine::GetConsoleInputHandle() { return m_inputHandle; }
HANDLE PrintMachine::GetConsoleOutputHandle() { return m_outputHandle; }
size_t PrintMachine::GetPrintSize() { return m_backBufferPrintSize; }
void PrintMachine::ResetBackBuffer() { memset(m_backBuffer.get(), 0, m_maxSize); }
void PrintMachine::FlagForBufferSwap() { m_bShouldSwapBuffer = true; }
void PrintMachine::SetPrintSize(const size_t n) { m_backBufferPrintSize = n; }
void PrintMachine::ResetConsolePointer() {}
