// Loads the compiled addon (build/Release/g16b200_napi.node).  UNTESTED here: no node in the build image.
module.exports = require("./build/Release/g16b200_napi.node");
