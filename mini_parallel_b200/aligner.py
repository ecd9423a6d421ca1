# placeholder until the host mirror lands
