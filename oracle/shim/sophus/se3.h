#pragma once
namespace Sophus { class SE3 {}; }
