{
  # node-gyp build of the N-API addon.  UNTESTED in this repository's image (no node / node-gyp):
  #   make -C ../shielded_pool_pinocchio_solana_b200/csrc     # builds libg16b200.so (nvcc, sm_100a)
  #   npm install && npx node-gyp rebuild
  "targets": [
    {
      "target_name": "g16b200_napi",
      "sources": ["g16b200_napi.cc"],
      "include_dirs": ["../include"],
      "defines": ["NAPI_VERSION=8"],
      "cflags_cc": ["-std=c++17", "-O2"],
      "libraries": [
        "-L<(module_root_dir)/../shielded_pool_pinocchio_solana_b200",
        "-lg16b200",
        "-Wl,-rpath,<(module_root_dir)/../shielded_pool_pinocchio_solana_b200"
      ]
    }
  ]
}
